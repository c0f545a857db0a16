#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for mode in 0 1; do
MOMA_BENCH_LOCKSTEP=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$mode bench.py --gpus 2 --quick --steps 60 > gpurun_out/k_bench_n2_$mode.json 2> gpurun_out/k_bench_n2_$mode.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/k_bench_n2_$mode.json"))
print("lockstep=$mode N=2 C3 weak ms/step", round(d['ms_per_step'],4), "dist", d['per_step_ms_rank0'], "e2e", round(d['e2e']['ms_per_step'],4), "nopdl", d.get('ms_per_step_without_pdl'))
PY
done
