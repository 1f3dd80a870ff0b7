# full ncu capture of one kernel of the C4 step: bash tools/run_ncu_one.sh <kernel regex> <out name>
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$1" -s 3 -c 1 -f -o gpurun_out/$2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --only-main > /dev/null 2>&1
ls -la gpurun_out/$2.ncu-rep
