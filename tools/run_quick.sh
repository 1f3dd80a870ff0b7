# quick check of a kernel change: GPU parity tests, then the main bench line (kernel_ms breakdown)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-main > gpurun_out/quick.json 2> gpurun_out/quick.err; tail -2 gpurun_out/quick.err
python - <<PY
import json
j=json.load(open('gpurun_out/quick.json'))
k=j['kernel_ms']
print('value',round(j['value']),'e2e',round(j['e2e']['value']), {a:round(v,4) for a,v in k.items() if a!='ekf_update_kernels'}, {a:round(v,4) for a,v in k['ekf_update_kernels'].items()}, 'frac', round(j['roofline']['frac'],4))
PY
