#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -8 gpurun_out/tests_gpu.log
timeout 300 python tools/step_time.py > gpurun_out/r2_step_time.log 2>&1; tail -9 gpurun_out/r2_step_time.log
timeout 300 python tools/profile_step.py 4 4 0 1 > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2_launches_w4a4_v2.csv python tools/profile_step.py 4 4 0 1 > gpurun_out/ncu1.log 2>&1
python tools/parse_launches.py gpurun_out/r2_launches_w4a4_v2.csv | cut -c1-120 | head -40
