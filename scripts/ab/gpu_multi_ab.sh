#!/usr/bin/env bash
# N-GPU A/B of the sharded step's exchange schedules (bench --quick only)
N=${1:-8}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29593 bench.py --gpus $N --quick --steps 100 > gpurun_out/ab_n${N}_$tag.json 2> gpurun_out/ab_n${N}_$tag.err; echo "bench $tag rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_n${N}_$tag.json"))
    print("$tag N=$N C3 weak ms/step", round(d['ms_per_step'],4), "dist", {k:round(v,4) for k,v in d['per_step_ms_rank0'].items()}, "launches", d['gpu_launches_per_step'], "parity", d['parity_check']['ok'])
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/ab_n${N}_$tag.err").read()[-1200:])
PY
}
run peer_raw MOMA_B200_PEER_GATHER_MAX_BYTES=16777216
run original MOMA_B200_PEER_GATHER_MAX_BYTES=67108864 MOMA_B200_GATHER_PROJECTIONS_MAX_BYTES=268435456 MOMA_B200_NCE_FUSED=never
run nccl_raw X=1
