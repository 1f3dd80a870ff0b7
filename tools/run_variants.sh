# experiment: bench kernel_ms for several values of an environment switch:  bash tools/run_variants.sh VAR v1 v2 ...
mkdir -p gpurun_out
VAR=$1; shift
for v in "$@"; do
env $VAR=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-main > gpurun_out/quick.json 2> gpurun_out/quick.err; tail -2 gpurun_out/quick.err
python - <<PY
import json
j=json.load(open('gpurun_out/quick.json'))
k=j['kernel_ms']
print('$VAR=$v value',round(j['value']),'e2e',round(j['e2e']['value']), {a:round(v,4) for a,v in k['ekf_update_kernels'].items()}, 'upd', round(k['ekf_update'],4))
PY
done
