#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 -p no:cacheprovider -k "gemm" > gpurun_out/tests_k.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_k.log
tail -3 gpurun_out/tests_k.log
timeout 600 python tools/k1_ab.py > gpurun_out/r2_k1_ab.txt 2>&1; cat gpurun_out/r2_k1_ab.txt | cut -c1-250
timeout 600 python tools/gemm_bench.py > gpurun_out/r2_gemm_bench.jsonl 2> gpurun_out/gemm_bench.err; tail -3 gpurun_out/gemm_bench.err; cat gpurun_out/r2_gemm_bench.jsonl | cut -c1-300
timeout 300 python tools/gemm_s8_bench.py > gpurun_out/r2_gemm_s8_bench.txt 2>&1; cat gpurun_out/r2_gemm_s8_bench.txt
timeout 300 python tools/step_time.py > gpurun_out/r2_step_time.log 2>&1; tail -9 gpurun_out/r2_step_time.log
