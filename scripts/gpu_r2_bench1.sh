#!/bin/bash
# round 2: bench lines at N=1 (default config 3 in full; configs 1, 2 in full; 4, 5 reduced as a functional check)
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1700 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name exit $?"; tail -2 gpurun_out/$name.err; cut -c1-900 gpurun_out/$name.json; }
run r2_bench_n1 --steps 4 --warmup 3
run r2_bench_cfg1_n1 --config 1 --steps 4 --warmup 3
run r2_bench_cfg2_n1 --config 2 --steps 2 --warmup 2 --no-cpu-baseline
run r2_bench_cfg4_n1_reduced --config 4 --epochs 2 --nsamples 16 --steps 2 --warmup 2 --no-cpu-baseline --e2e-blocks 2
run r2_bench_cfg5_n1_reduced --config 5 --epochs 2 --nsamples 8 --steps 2 --warmup 2 --no-cpu-baseline --e2e-blocks 2
