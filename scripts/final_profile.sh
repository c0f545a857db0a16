#!/bin/bash
# One pass over everything the round's profiles/ are built from (single GPU).  Outputs land in gpurun_out/.
set -x
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; tail -2 gpurun_out/final_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
timeout 300 python bench.py > gpurun_out/final_bench_c2.json 2> gpurun_out/final_bench_c2.err || exit 1
timeout 300 python bench.py --config C3 --no-cpu-baseline > gpurun_out/final_bench_c3.json 2> gpurun_out/final_bench_c3.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
# launch list of the same command (cold-cache, serialised: compare SHARES with the timeline, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/final_launches_c2.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_launches.log 2>&1
# full captures of the two roofline kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ema_multi_kernel -s 4 -c 1 -o gpurun_out/final_ema \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_ema.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nce_tc2_kernel -s 4 -c 1 -o gpurun_out/final_nce \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_nce.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm3xtf32 -s 12 -c 1 -o gpurun_out/final_gemm \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_gemm.log 2>&1
timeout 200 python scripts/profile_step.py C2 ovl flush > gpurun_out/final_timeline_c2.txt 2>&1
timeout 200 python scripts/sweep_nce.py > gpurun_out/final_nce_sweep.txt 2>&1
timeout 100 python scripts/check_linear.py > gpurun_out/final_linear.txt 2>&1
ls -la gpurun_out/final_*
