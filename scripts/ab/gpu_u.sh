#!/usr/bin/env bash
# head_dim 64 on the tensor-core attention kernels: tests + C5 with / without
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -x -q -k "attention or attn or mocoatt" > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/u_pytest.log
for v in 1 0; do
  MOMA_B200_ATTN_TC64=$v timeout 300 python bench.py --config C5 --quick --no-cpu-baseline --steps 60 > gpurun_out/u_c5_tc64_$v.json 2> gpurun_out/u_c5_tc64_$v.err; echo "bench rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/u_c5_tc64_$v.json'))
print('tc64=$v C5 ms/step', round(d['ms_per_step'],4), 'parity', d['parity_check']['ok'])
for k,x in sorted(d['kernel_shares']['families'].items(), key=lambda kv:-kv[1]['us'])[:8]: print(f\"{x['us']:8.1f} us x{x['launches']:<5} {k}\")
"
done
