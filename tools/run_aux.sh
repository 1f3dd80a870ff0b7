# GPU tests of the search/detector/particle rows + the aux bench numbers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py -x -q 2>&1 | tail -4
timeout 300 python - <<PY
import json, bench
print(json.dumps(bench.aux_rates(), indent=1))
PY
