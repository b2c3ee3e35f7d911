#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py 4 4 0 1 > gpurun_out/prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_ -c 4 -o gpurun_out/r2_attn_prof -f python tools/profile_step.py 4 4 0 1 > gpurun_out/ncu_attn.log 2>&1
tail -5 gpurun_out/ncu_attn.log
ls -la gpurun_out/*.ncu-rep
