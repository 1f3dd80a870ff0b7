mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "H2 |passed|failed|Error|error|assert" | tail -15 > gpurun_out/r2_t10.log; cat gpurun_out/r2_t10.log
run() { env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $EXTRA > gpurun_out/r2_b10.json 2> gpurun_out/r2_b10.err; python -c "
import json,sys; j=json.load(open('gpurun_out/r2_b10.json')); k=j['kernel_ms']; print(sys.argv[1:], round(j['value']), round(j['e2e']['value']), round(k['ekf_update'],4), {a:round(b,4) for a,b in k['ekf_update_kernels'].items()})" "$@"; }
run A=1
EXTRA="--step-groups 2" run A=2
