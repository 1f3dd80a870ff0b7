# The default bench line again (after a change that does not touch the kernels, e.g. the traffic table): bench + clocks
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r02_clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -2 gpurun_out/r02_bench.err
kill $SMI
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench.err
python - <<PY
import json
j=json.load(open('gpurun_out/r02_bench.json'))
print('value', round(j['value']), 'e2e', round(j['e2e']['value']), 'frac', round(j['roofline']['frac'],4), 'traffic', j['roofline']['traffic'])
r=json.load(open('gpurun_out/r02_bench_reference_arm.json'))
print('reference arm', {k:r.get(k) for k in ('impl','value','unit','steps','ms_per_step')}, r.get('cpu_baseline',{}).get('cores'))
PY
