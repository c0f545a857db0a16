#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py tests/test_sharded_gpu.py -q --timeout 300 > gpurun_out/n_pytest.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/n_pytest.log
run() { tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --quick --steps 80 > gpurun_out/n_bench_$tag.json 2> gpurun_out/n_bench_$tag.err; echo "bench $tag rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/n_bench_$tag.json"))
    print("$tag N=2 C3 weak ms/step", round(d['ms_per_step'],4), "dist", {k:round(v,4) for k,v in d['per_step_ms_rank0'].items()}, "e2e", round(d['e2e']['ms_per_step'],4), "launches", d['gpu_launches_per_step'], "parity", d['parity_check']['ok'], d['parity_check']['graph_vs_sequential_replicated']['dfeat_s_rel'])
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/n_bench_$tag.err").read()[-1500:])
PY
}
run default X=1
run forced MOMA_B200_PEER_GATHER_MAX_BYTES=1000 MOMA_B200_GATHER_PROJECTIONS_MAX_BYTES=1000 MOMA_B200_NCE_FUSED=1
