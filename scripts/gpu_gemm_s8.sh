#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "s8" 2>&1 | tail -6
timeout 300 python tools/gemm_s8_bench.py > gpurun_out/r2_gemm_s8_bench_2cta.txt 2>&1; cat gpurun_out/r2_gemm_s8_bench_2cta.txt | tail -8
