# GPU parity tests with the default knobs and with every scheduling knob on, then the knob sweep (tools/sweep_tuning.py)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
SL2_TUNE="0=7000,1=3000,2=1,3=1,4=1" timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/sweep_tuning.py 2>&1 | tail -40
