#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 -p no:cacheprovider -k "attention" > gpurun_out/tests_j.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_j.log
tail -3 gpurun_out/tests_j.log
timeout 300 python tools/profile_step.py 4 4 0 1 > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2_launches_w4a4_v5.csv python tools/profile_step.py 4 4 0 1 > gpurun_out/ncu1.log 2>&1
python tools/parse_launches.py gpurun_out/r2_launches_w4a4_v5.csv | cut -c1-120 | grep -i "attn\|total"
