#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_modules_gpu.py tests/test_dp_gpu.py -q -m gpu --timeout 600 -p no:cacheprovider -k "sym or 16bit or full_shape or smooth or dp or two_engine" > gpurun_out/tests_b.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_b.log
tail -12 gpurun_out/tests_b.log
timeout 600 python tools/e2e_chain_debug.py 4 20 128 > gpurun_out/chain_debug.log 2>&1; echo "chain exit $?" >> gpurun_out/chain_debug.log
grep -n "oq debug\|FAILED\|iter 0 \|iter 19\|exit" gpurun_out/chain_debug.log | head -40
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -p no:cacheprovider -k "not large and not full_size and not llama_shapes" > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r2_sanitizer_memcheck.log
tail -8 gpurun_out/r2_sanitizer_memcheck.log
