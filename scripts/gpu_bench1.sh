#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python bench.py --steps 4 --warmup 3 > gpurun_out/r2_bench_n1_v6.json 2> gpurun_out/r2_bench_n1_v6.err; echo "exit $?"; tail -2 gpurun_out/r2_bench_n1_v6.err; cut -c1-400 gpurun_out/r2_bench_n1_v6.json
