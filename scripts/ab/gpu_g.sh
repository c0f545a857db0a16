#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_round2_gpu.py tests/test_tc_gpu.py -q --timeout 300 -k "fused or tc_gpu or overlapped" > gpurun_out/g_pytest_fused.log 2>&1; echo "fused tests rc=$?"; tail -5 gpurun_out/g_pytest_fused.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/g_pytest.log
timeout 900 python bench.py > gpurun_out/g_bench_n1.json 2> gpurun_out/g_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/g_bench_n1.err
timeout 300 python scripts/profile_step.py C3 ovl flush > gpurun_out/g_timeline_c3.txt 2>&1; echo "timeline rc=$?"
timeout 300 python scripts/profile_step.py C2 ovl flush > gpurun_out/g_timeline_c2.txt 2>&1
