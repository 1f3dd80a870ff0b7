mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t11.log; cat gpurun_out/r2_t11.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b11.json 2> gpurun_out/r2_b11.err; tail -3 gpurun_out/r2_b11.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_b11.json'))
print('value',round(j['value']),'e2e',round(j['e2e']['value']),'frac',round(j['roofline']['frac'],3), j['kernel_ms'])
for k,v in j['configs'].items():
    if 'value' in v: print(k, round(v['value']), round(v.get('e2e',{}).get('value',0)), v.get('kernel_ms') or v)
    else: print(k, v)
print(j.get('cpu_baseline'))
PY
