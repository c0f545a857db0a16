#!/usr/bin/env bash
# N-GPU pass: bench (quick) + timeline of the sharded C3 step;  usage: gpu_multi.sh N
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --quick --steps 80 > gpurun_out/m_bench_n$N.json 2> gpurun_out/m_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/m_bench_n$N.json"))
print("N=$N C3 weak ms/step", round(d['ms_per_step'],4), "value", round(d['value']), "dist", d['per_step_ms_rank0'], "e2e", round(d['e2e']['ms_per_step'],4), "launches", d['gpu_launches_per_step'], "parity", d['parity_check']['ok'], d.get('exchange'))
for k,v in sorted(d['kernel_shares']['families'].items(), key=lambda kv:-kv[1]['us']): print(f"{v['us']:8.1f} us x{v['launches']:<5} {k}")
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29582 scripts/profile_step.py C3 ovl flush > gpurun_out/m_timeline_c3_n$N.txt 2> gpurun_out/m_timeline_c3_n$N.err; echo "timeline rc=$?"
grep -v "^$" gpurun_out/m_timeline_c3_n$N.txt | grep -v Warn | head -64 | cut -c1-135
