mkdir -p gpurun_out
for g in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-main --step-groups $g > gpurun_out/quick.json 2> gpurun_out/quick.err; tail -2 gpurun_out/quick.err
python - <<PY
import json
j=json.load(open('gpurun_out/quick.json'))
print('groups=$g value',round(j['value']),'e2e',round(j['e2e']['value']),'ms',round(j['ms_per_step'],4))
PY
done
