#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "gemm" 2>&1 | tail -12
timeout 300 python tools/gemm_bench.py > gpurun_out/r2_gemm_bench_2cta.jsonl 2> gpurun_out/gemm_bench.err; tail -3 gpurun_out/gemm_bench.err; cat gpurun_out/r2_gemm_bench_2cta.jsonl
