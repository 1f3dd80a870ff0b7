# One gpurun call: tests, the default bench line, the ncu launch list of the same command and full captures of
# the update kernels and the search kernel (C4 and C3).  Outputs under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r02_clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -2 gpurun_out/r02_bench.err
kill $SMI
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --only-main"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv $B > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"upd_|search_kernel|predict_kernel|cull_kernel" -s 9 -c 9 -f -o gpurun_out/r02_step_C4 $B > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 1 -c 1 -f -o gpurun_out/r02_search_C3 $B --config C3 > /dev/null 2>&1
ls -la gpurun_out | grep r02
