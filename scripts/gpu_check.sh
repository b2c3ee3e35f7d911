#!/bin/bash
# One gpurun call: kernel parity tests, then (if green) the GEMM microbenchmark.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 300 -p no:cacheprovider "$@" > gpurun_out/kernels.log 2>&1
echo "pytest exit $?" >> gpurun_out/kernels.log
tail -40 gpurun_out/kernels.log
timeout 600 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1
echo "gemm_bench exit $?" >> gpurun_out/gemm_bench.log
tail -12 gpurun_out/gemm_bench.log
