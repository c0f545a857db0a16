#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python scripts/probe_launch.py > gpurun_out/c_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/c_probe.txt
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_reference_loop_gpu.py -q --timeout 600 > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/c_pytest.log
