#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -q --timeout 300 > gpurun_out/i_pytest.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/i_pytest.log
for cfg in C2 C3; do for ov in 1 2 3; do
  MOMA_B200_GEMM_OVERSUB=$ov timeout 300 python bench.py --config $cfg --quick --no-cpu-baseline --steps 60 > gpurun_out/i_${cfg}_ov${ov}.json 2> gpurun_out/i_${cfg}_ov${ov}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/i_${cfg}_ov${ov}.json"))
f=d['kernel_shares']['families']
print("$cfg oversub=$ov  ms/step", round(d['ms_per_step'],4), "launches", d['gpu_launches_per_step'], "gemm us", f['gemm3xtf32']['us'], "other", f.get('other'), "parity", d['parity_check']['ok'])
PY
done; done
timeout 300 python scripts/profile_step.py C3 ovl flush > gpurun_out/i_timeline_c3.txt 2>&1
