#!/usr/bin/env bash
# tcgen05 GEMM: probe against the warp-level kernel, tests, and the N=1 step with the dispatch on / off
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
MOMA_B200_GEMM_TC=0 python scripts/probe_gemm_tc.py > gpurun_out/s_probe_gemm.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/s_probe_gemm.txt | tail -32
python -m pytest tests/test_gemm_tc_gpu.py -x -q > gpurun_out/s_pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -4 gpurun_out/s_pytest_gemm.log
for tc in 1 0; do
  MOMA_B200_GEMM_TC=$tc python bench.py --quick --steps 100 > gpurun_out/s_bench_tc$tc.json 2> gpurun_out/s_bench_tc$tc.err; echo "bench tc=$tc rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/s_bench_tc$tc.json"))
print("tc=$tc C3 ms/step", round(d['ms_per_step'],4), "e2e", d['e2e']['value'], "launches", d['gpu_launches_per_step'], "parity", d['parity_check'])
PY
done
