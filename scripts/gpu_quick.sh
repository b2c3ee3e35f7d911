#!/bin/bash
# Quick GPU check: all -m gpu tests, step timing, launch list of one W4A4 step.  Logs -> gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -15 gpurun_out/tests_gpu.log
timeout 300 python tools/step_time.py > gpurun_out/step_time.log 2>&1; tail -5 gpurun_out/step_time.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_w4a4.csv python tools/profile_step.py 4 4 0 1 > gpurun_out/ncu1.log 2>&1
python tools/parse_launches.py gpurun_out/launches_w4a4.csv | cut -c1-110 | head -40
