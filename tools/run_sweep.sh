mkdir -p gpurun_out
for b in 296 444 592 888; do timeout 300 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --only-main --streams $b > gpurun_out/r2_b14.json 2> gpurun_out/r2_b14.err; python - <<PY
import json
j=json.load(open('gpurun_out/r2_b14.json'))
k=j['kernel_ms']
print($b,'value',round(j['value']),'e2e',round(j['e2e']['value']), {a:round(v,4) for a,v in k.items() if a!='ekf_update_kernels'}, {a:round(v,4) for a,v in k['ekf_update_kernels'].items()})
PY
done; tail -2 gpurun_out/r2_b14.err
