#!/usr/bin/env bash
# 2-GPU pass: sharded tests, bench at N=2 (parity self-check at every config), v3 sweep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_gpu.py tests/test_reference_loop_gpu.py tests/test_tc_gpu.py -q --timeout 300 > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/e_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 > gpurun_out/e_bench_n2.json 2> gpurun_out/e_bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/e_bench_n2.err
timeout 300 python scripts/sweep_nce.py > gpurun_out/e_sweep.txt 2>&1; cat gpurun_out/e_sweep.txt
head -c 600 gpurun_out/e_bench_n2.json
