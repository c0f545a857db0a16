#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_round2_gpu.py tests/test_tc_gpu.py -q --timeout 300 -k "fused or tc_gpu or overlapped" > gpurun_out/h_pytest_fused.log 2>&1; echo "fused tests rc=$?"; tail -3 gpurun_out/h_pytest_fused.log
for cfg in C2 C3; do for fused in 1 0; do for ema in late early; do
  MOMA_B200_NCE_FUSED=$fused MOMA_B200_EMA_FORK=$ema timeout 300 python bench.py --config $cfg --quick --no-cpu-baseline --steps 60 > gpurun_out/h_${cfg}_f${fused}_${ema}.json 2> gpurun_out/h_${cfg}_f${fused}_${ema}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/h_${cfg}_f${fused}_${ema}.json"))
n=d['roofline_north_star']
print("$cfg fused=$fused ema=$ema  ms/step", round(d['ms_per_step'],4), "launches", d['gpu_launches_per_step'], "nce in-step us", n['us_in_step'], "nce/launch", round(n['us_per_launch'],2), "parity", d['parity_check']['ok'])
PY
done; done; done
timeout 300 python scripts/profile_step.py C2 ovl flush > gpurun_out/h_timeline_c2.txt 2>&1
timeout 300 python scripts/profile_step.py C3 ovl flush > gpurun_out/h_timeline_c3.txt 2>&1
