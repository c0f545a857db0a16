#!/usr/bin/env bash
# Round-2 evidence pass (single GPU): tests, smoke, bench (both arms), ncu launch list + full captures, timelines, sweeps.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/p_clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/p_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/p_bench_n1.json 2> gpurun_out/p_bench_n1.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/p_bench_ref.json 2> gpurun_out/p_bench_ref.err; echo "ref rc=$?"
kill $SMI
timeout 300 python scripts/profile_step.py C3 ovl flush > gpurun_out/p_timeline_c3.txt 2>&1
timeout 300 python scripts/profile_step.py C2 ovl flush > gpurun_out/p_timeline_c2.txt 2>&1
timeout 300 python scripts/sweep_nce.py > gpurun_out/p_sweep_nce.txt 2>&1
timeout 200 python scripts/probe_launch.py > gpurun_out/p_probe_launch.txt 2>&1
MOMA_B200_LIB=moma_b200/lib/libmoma_b200_ablate.so timeout 200 python scripts/trace_nce_life.py > gpurun_out/p_nce_life.txt 2>&1
# ncu: launch list of the step command, then full captures of the top kernels (the plain run of the same command first)
timeout 300 python scripts/run_step_once.py C3 6 > gpurun_out/p_plain_c3.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/p_launches_c3.csv \
    python scripts/run_step_once.py C3 6 > gpurun_out/p_ncu_launches_c3.log 2>&1; echo "ncu launches rc=$?"
for k in nce_tc3_kernel attn_fwd_tc_kernel attn_bwd_dkv_tc_kernel gemm3xtf32_kernel ema_multi_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/p_full_$k \
      python scripts/run_step_once.py C3 6 > gpurun_out/p_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 300 python scripts/run_step_once.py C2 6 > gpurun_out/p_plain_c2.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/p_launches_c2.csv \
    python scripts/run_step_once.py C2 6 > gpurun_out/p_ncu_launches_c2.log 2>&1
ls -la gpurun_out/p_*
