# experiment: the main bench line for several prebuilt library variants (scenelib2_b200/_variants/*.so)
mkdir -p gpurun_out
for v in scenelib2_b200/_variants/*.so; do
cp $v scenelib2_b200/libsl2b200.so
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-main > gpurun_out/quick.json 2> gpurun_out/quick.err; tail -2 gpurun_out/quick.err
python - <<PY
import json
j=json.load(open('gpurun_out/quick.json'))
k=j['kernel_ms']
print('$v value',round(j['value']), {a:round(v,4) for a,v in k['ekf_update_kernels'].items()}, 'upd', round(k['ekf_update'],4))
PY
done
