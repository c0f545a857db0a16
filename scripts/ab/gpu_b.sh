#!/usr/bin/env bash
# round-2 GPU pass B: v3 InfoNCE kernel -- parity tests first, then trace / sweep (MUFU vs polynomial exp2), then bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py -q --timeout 300 > gpurun_out/b_pytest_tc.log 2>&1; echo "tc tests rc=$?"; tail -3 gpurun_out/b_pytest_tc.log
MOMA_B200_LIB=moma_b200/lib/libmoma_b200_ablate.so timeout 300 python scripts/trace_nce_life.py > gpurun_out/b_nce_life.txt 2>&1; echo "life rc=$?"
timeout 300 python scripts/sweep_nce.py > gpurun_out/b_sweep_mufu.txt 2>&1; echo "sweep rc=$?"
MOMA_B200_NCE_POLY=1 timeout 300 python scripts/sweep_nce.py > gpurun_out/b_sweep_poly.txt 2>&1; echo "sweep poly rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/b_pytest.log
timeout 900 python bench.py > gpurun_out/b_bench_n1.json 2> gpurun_out/b_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/b_bench_n1.err
cat gpurun_out/b_sweep_mufu.txt gpurun_out/b_sweep_poly.txt
