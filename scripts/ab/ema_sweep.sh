for r in 1 2 3 4 8; do
  MOMA_B200_EMA_CTAS_PER_SM=$r timeout 200 python bench.py --steps 100 --warmup 5 > gpurun_out/b_ema$r.json 2> gpurun_out/b_ema$r.err
  python -c "
import json; d=json.load(open('gpurun_out/b_ema$r.json')); print('ctas/sm', $r, 'ms/step', round(d['ms_per_step'],4), 'value', int(d['value']), 'e2e', int(d['e2e']['value']), 'ema us', round(d['roofline']['us_per_launch'],1), 'frac', round(d['roofline']['frac'],3))"
done
