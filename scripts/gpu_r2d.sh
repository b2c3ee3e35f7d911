#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_kernels_gpu.py -q -m gpu --timeout 600 -p no:cacheprovider -k "opt_w6a6 or s8" > gpurun_out/tests_d.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_d.log
tail -6 gpurun_out/tests_d.log
OQ_NO_GRAPH=1 timeout 900 python tools/e2e_chain_debug.py 3 20 128 > gpurun_out/chain_debug2.log 2>&1; echo "chain exit $?" >> gpurun_out/chain_debug2.log
grep -v "iter [0-9]* loss" gpurun_out/chain_debug2.log | head -120
