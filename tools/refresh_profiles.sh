# after tools/capture_profiles.sh: gpurun_out/ -> profiles/ (summaries, traffic table, bench line, logs)
set -e
python tools/ncu_summary.py gpurun_out/r02_step_C4.ncu-rep > profiles/r02_step_C4_ncu.txt 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02_search_C3.ncu-rep > profiles/r02_search_C3_ncu.txt 2>/dev/null
python tools/ncu_traffic.py C4=gpurun_out/r02_step_C4.ncu-rep C3=gpurun_out/r02_search_C3.ncu-rep > profiles/r02_dram_bytes.json
cp gpurun_out/r02_bench.json profiles/r02_bench_default.json
cp gpurun_out/r02_launches.csv profiles/r02_launches.csv
cp gpurun_out/r02_clocks.csv profiles/r02_clocks_during_bench.csv
cp gpurun_out/r02_pytest_gpu.log profiles/r02_pytest_gpu.log
