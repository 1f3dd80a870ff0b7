#!/bin/bash
# Round-end validation on the GPU box: tests, smoke, benches, launch list, ncu captures, sanitizer.
set -x
O=gpurun_out/final; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 400 python bench.py > $O/bench_C4.json 2> $O/bench_C4.err
for c in C1 C2 C3; do timeout 300 python bench.py --config $c --no-cpu-baseline > $O/bench_$c.json 2>> $O/bench_C4.err; done
timeout 300 python bench.py --streams 1 --no-cpu-baseline > $O/bench_C4_1stream.json 2>> $O/bench_C4.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2>> $O/bench_C4.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:update_kernel -s 4 -c 1 -f -o $O/upd python bench.py --steps 2 --warmup 2 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:search_kernel -s 4 -c 1 -f -o $O/srch python bench.py --steps 2 --warmup 2 --no-cpu-baseline > /dev/null 2>&1
cp scenelib2_b200/libsl2b200.so $O/lib_final.so
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log; tail -4 $O/memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize.py > $O/racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/racecheck.log; tail -4 $O/racecheck.log
ls -la $O
