#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/f_pytest.log
cat gpurun_out/loop_parity_diag.txt
timeout 900 python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/f_bench_n1.err
timeout 300 python scripts/profile_step.py C3 ovl flush > gpurun_out/f_timeline_c3.txt 2>&1; echo "timeline rc=$?"
timeout 300 python scripts/profile_step.py C2 ovl flush > gpurun_out/f_timeline_c2.txt 2>&1
