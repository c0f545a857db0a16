#!/usr/bin/env bash
# Round-2 closing evidence pass (single GPU): tests, smoke, bench (both arms), timeline, GEMM probe, ncu launch list + tcgen05 GEMM capture
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "ref rc=$?"
timeout 300 python scripts/profile_step.py C3 ovl flush > gpurun_out/f_timeline_c3.txt 2>&1
MOMA_B200_GEMM_TC=0 timeout 300 python scripts/probe_gemm_tc.py > gpurun_out/f_gemm_tc_probe.txt 2>&1
timeout 300 python scripts/run_step_once.py C3 6 > gpurun_out/f_plain_c3.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/f_launches_c3.csv \
    python scripts/run_step_once.py C3 6 > gpurun_out/f_ncu_launches_c3.log 2>&1; echo "ncu launches rc=$?"
for k in gemm_tc_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/f_full_$k \
      python scripts/run_step_once.py C3 6 > gpurun_out/f_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
python - <<'PY'
import json
d=json.load(open("gpurun_out/f_bench_n1.json"))
print("C3 ms/step", round(d['ms_per_step'],4), "value", round(d['value']), "e2e", d['e2e'].get('value'), "launches", d['gpu_launches_per_step'], "parity", d['parity_check']['ok'])
print("roofline", d['roofline'])
for k,v in sorted(d['kernel_shares']['families'].items(), key=lambda kv:-kv[1]['us']): print(f"{v['us']:8.1f} us x{v['launches']:<5} {k}")
PY
