#!/bin/bash
# Round-end validation on the GPU box: tests, smoke, benches (ncu captures / sanitizer: see profiles/).
set -x
O=gpurun_out/final2; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 400 python bench.py > $O/bench_C4.json 2> $O/bench_C4.err; tail -c 400 $O/bench_C4.json
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log; tail -3 $O/memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize.py > $O/racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/racecheck.log; tail -3 $O/racecheck.log
