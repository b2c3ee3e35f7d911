#!/bin/bash
# ncu --set full of the weight-quantiser kernels of one W4A4 LLaMA-7B step (after the plain run exited 0)
mkdir -p gpurun_out
K=${1:-wq_}
timeout 300 python tools/profile_step.py 4 4 0 1 > gpurun_out/prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$K -c 8 -o gpurun_out/r2_wq_prof2 -f python tools/profile_step.py 4 4 0 1 > gpurun_out/ncu_wq.log 2>&1
tail -3 gpurun_out/ncu_wq.log
