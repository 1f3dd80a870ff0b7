# PDL with the trigger before / after the wait, against PDL off, at 1 / 16 / 296 streams (C4 main leg).  Needs the two
# builds side by side: scenelib2_b200/_variants/wait_first.so (the tree as it is) and trigger_first.so (pdl_prologue
# with the two instructions swapped).
mkdir -p gpurun_out
for v in wait_first trigger_first; do cp scenelib2_b200/_variants/$v.so scenelib2_b200/libsl2b200.so
for b in 1 16 296; do for m in 0 1; do
if [ $v = trigger_first ] && [ $m = 0 ]; then continue; fi
SL2_TUNE="0=$m" timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --only-main --streams $b > gpurun_out/q.json 2> gpurun_out/q.err || tail -2 gpurun_out/q.err
python - <<PY
import json
j=json.load(open('gpurun_out/q.json'))
print('$v streams $b pdl $m  ms/step %.4f  frames/s %d  e2e %d' % (j['ms_per_step'], j['value'], j['e2e']['value']))
PY
done; done; done
cp scenelib2_b200/_variants/wait_first.so scenelib2_b200/libsl2b200.so
SL2_TUNE="0=1" timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
