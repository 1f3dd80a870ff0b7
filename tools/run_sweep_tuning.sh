# GPU parity tests with the default knobs and with every scheduling knob on, then the knob sweep (tools/sweep_tuning.py)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
SL2_TUNE="0=1,1=0" timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/sweep_tuning.py 2>&1 | tail -40
