mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t13.log; cat gpurun_out/r2_t13.log
for c in C4 C3 C2 C1; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-main --config $c > gpurun_out/r2_b13_$c.json 2> gpurun_out/r2_b13.err; python - <<PY
import json
j=json.load(open('gpurun_out/r2_b13_$c.json'))
print('$c value',round(j['value']),'e2e',round(j['e2e']['value']), {k:round(v,4) for k,v in j['kernel_ms'].items() if k!='ekf_update_kernels'})
PY
done; tail -2 gpurun_out/r2_b13.err
