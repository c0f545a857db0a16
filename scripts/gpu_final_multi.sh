#!/usr/bin/env bash
# Closing N-GPU pass: the DEFAULT bench command of the driver (all blocks), optionally the sharded test and the timeline
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
if [ "${2:-}" = "test" ]; then
  timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q > gpurun_out/g_pytest_n$N.log 2>&1; echo "pytest sharded rc=$?"; tail -3 gpurun_out/g_pytest_n$N.log
fi
S=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N > gpurun_out/g_bench_n$N.json 2> gpurun_out/g_bench_n$N.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/g_bench_n$N.json"))
    print("N=$N C3 weak ms/step", round(d['ms_per_step'],4), "value", round(d['value']), "dist", {k:round(v,4) for k,v in d['per_step_ms_rank0'].items()}, "e2e", d['e2e'].get('value'), "launches", d['gpu_launches_per_step'], "parity", d['parity_check']['ok'], d.get('exchange'))
    for k,v in (d.get('other_configs') or {}).items():
        print("  extra", k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms_per_step','value','scaling','parity_ok')})
    for k,v in sorted(d['kernel_shares']['families'].items(), key=lambda kv:-kv[1]['us'])[:12]: print(f"{v['us']:8.1f} us x{v['launches']:<5} {k}")
except Exception as e:
    print("failed", e); print(open("gpurun_out/g_bench_n$N.err").read()[-2500:])
PY
if [ "${3:-}" = "timeline" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29592 scripts/profile_step.py C3 ovl flush > gpurun_out/g_timeline_c3_n$N.txt 2> gpurun_out/g_timeline_c3_n$N.err; echo "timeline rc=$?"
  grep -v "^$" gpurun_out/g_timeline_c3_n$N.txt | grep -v Warn | head -60 | cut -c1-135
fi
