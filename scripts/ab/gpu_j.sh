#!/usr/bin/env bash
# 2-GPU: timeline of the sharded C3 step + bench quick
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 scripts/profile_step.py C3 ovl flush > gpurun_out/j_timeline_c3_n2.txt 2> gpurun_out/j_timeline_c3_n2.err; echo "timeline rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --quick --steps 60 > gpurun_out/j_bench_n2.json 2> gpurun_out/j_bench_n2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/j_bench_n2.json"))
print("N=2 C3 weak ms/step", d['ms_per_step'], "e2e", d['e2e']['ms_per_step'], "launches", d['gpu_launches_per_step'], d['parity_check']['ok'])
for k,v in sorted(d['kernel_shares']['families'].items(), key=lambda kv:-kv[1]['us']): print(f"{v['us']:8.1f} us x{v['launches']:<5} {k}")
PY
grep -v "^$" gpurun_out/j_timeline_c3_n2.txt | head -70 | cut -c1-140
