#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py -q --timeout 300 > gpurun_out/d_pytest_tc.log 2>&1; echo "tc tests rc=$?"; tail -3 gpurun_out/d_pytest_tc.log
MOMA_B200_LIB=moma_b200/lib/libmoma_b200_ablate.so timeout 300 python scripts/trace_nce_life.py > gpurun_out/d_nce_life.txt 2>&1; echo "life rc=$?"; head -30 gpurun_out/d_nce_life.txt
timeout 300 python scripts/sweep_nce.py > gpurun_out/d_sweep_mufu.txt 2>&1; echo "sweep rc=$?"
MOMA_B200_NCE_POLY=1 timeout 300 python scripts/sweep_nce.py > gpurun_out/d_sweep_poly.txt 2>&1; echo "sweep poly rc=$?"
cat gpurun_out/d_sweep_mufu.txt gpurun_out/d_sweep_poly.txt
timeout 900 python -m pytest tests/test_reference_loop_gpu.py tests/test_parity_gpu.py -q --timeout 600 > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/d_pytest.log
