#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_round2_gpu.py tests/test_parity_gpu.py -q --timeout 300 > gpurun_out/l_pytest.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/l_pytest.log
for cfg in C2 C3; do for ov in 2 4; do
  MOMA_B200_GEMM_OVERSUB=$ov timeout 300 python bench.py --config $cfg --quick --no-cpu-baseline --steps 80 > gpurun_out/l_${cfg}_ov${ov}.json 2> gpurun_out/l_${cfg}_ov${ov}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/l_${cfg}_ov${ov}.json"))
f=d['kernel_shares']['families']; n=d['roofline_north_star']
print("$cfg oversub=$ov  ms/step", round(d['ms_per_step'],4), "launches", d['gpu_launches_per_step'], "gemm us", f['gemm3xtf32']['us'], "nce/launch", round(n['us_per_launch'],2), "in-step", n['us_in_step'], "parity", d['parity_check']['ok'])
PY
done; done
timeout 300 python scripts/run_step_once.py C3 6 > gpurun_out/l_plain_c3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nce_tc3_kernel -s 3 -c 1 -o gpurun_out/l_full_nce_tc3_unfused \
      python scripts/run_step_once.py C3 6 > gpurun_out/l_ncu_nce.log 2>&1; echo "ncu rc=$?"
