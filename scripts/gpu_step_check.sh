#!/bin/bash
# kernel + engine tests, then the serialised launch list of one W4A4 step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -m gpu -x -k "${1:-swiglu or block_calibration or emulation}" 2>&1 | tail -5
timeout 300 python tools/profile_step.py 4 4 0 1 > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_w4a4_v8.csv python tools/profile_step.py 4 4 0 1 > gpurun_out/ncu_lean.log 2>&1
python tools/parse_launches.py gpurun_out/r2_launches_w4a4_v8.csv 2>&1 | head -34
