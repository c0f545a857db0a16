#!/usr/bin/env bash
# round-2 GPU pass A: smoke, full GPU test suite, default bench (both arms), life-cycle trace of the InfoNCE kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/a_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/a_pytest.log
timeout 900 python bench.py > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/a_bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; echo "ref rc=$?"
MOMA_B200_LIB=moma_b200/lib/libmoma_b200_ablate.so timeout 300 python scripts/trace_nce_life.py > gpurun_out/a_nce_life.txt 2>&1; echo "life rc=$?"
head -c 1500 gpurun_out/a_bench_n1.json
