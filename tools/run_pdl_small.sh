# PDL on / off against the number of camera streams of the batch (C4, frames resident): where does it pay?
mkdir -p gpurun_out
for b in 1 4 16 37 74 148; do for m in 0 1; do
SL2_TUNE="0=$m" timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --only-main --streams $b > gpurun_out/q.json 2> gpurun_out/q.err || tail -2 gpurun_out/q.err
python - <<PY
import json
j=json.load(open('gpurun_out/q.json'))
print('streams $b pdl $m  ms/step %.4f  frames/s %d  e2e %d' % (j['ms_per_step'], j['value'], j['e2e']['value']))
PY
done; done
