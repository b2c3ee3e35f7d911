#!/bin/bash
# round 2: head-dim-64 fused attention: parity tests, then configs 1 and 5 (reduced) bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" -s 2>&1 | tail -25
run() { name=$1; shift; timeout 1200 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name exit $?"; tail -2 gpurun_out/$name.err; cut -c1-1200 gpurun_out/$name.json; }
run r2_bench_cfg1_n1_v3 --config 1 --steps 4 --warmup 3
run r2_bench_cfg5_n1_reduced_v3 --config 5 --epochs 2 --nsamples 8 --steps 2 --warmup 2 --no-cpu-baseline --e2e-blocks 2
