#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -m gpu --timeout 600 -p no:cacheprovider -k "attention or gemm_fq or emulation or full_shape" > gpurun_out/tests_i.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_i.log
tail -5 gpurun_out/tests_i.log
timeout 600 python tools/k1_ab.py > gpurun_out/r2_k1_ab.txt 2>&1; cat gpurun_out/r2_k1_ab.txt | cut -c1-260
timeout 300 python tools/step_time.py > gpurun_out/r2_step_time.log 2>&1; tail -9 gpurun_out/r2_step_time.log
timeout 300 python tools/profile_step.py 4 4 0 1 > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2_launches_w4a4_v4.csv python tools/profile_step.py 4 4 0 1 > gpurun_out/ncu1.log 2>&1
python tools/parse_launches.py gpurun_out/r2_launches_w4a4_v4.csv | cut -c1-120 | head -16
