#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "calibration_variant" 2>&1 | tail -15
T=${1:-256}
OQ_WQ_LEAN_THREADS=$T timeout 300 python tools/profile_step.py 4 4 0 1 > gpurun_out/prof_plain.log 2>&1 && \
OQ_WQ_LEAN_THREADS=$T timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_w4a4_lean$T.csv python tools/profile_step.py 4 4 0 1 > gpurun_out/ncu_lean.log 2>&1
python tools/parse_launches.py gpurun_out/r2_launches_w4a4_lean$T.csv 2>&1 | head -40
