#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 -p no:cacheprovider -k "fused_attention" -s > gpurun_out/tests_e.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_e.log
grep -n "fused attention\|passed\|failed\|Error\|error\|exit" gpurun_out/tests_e.log | head -40
OQ_NO_GRAPH=1 OQ_FUSED_ATTN=0 timeout 900 python tools/e2e_chain_debug.py 3 20 128 > gpurun_out/chain_debug3.log 2>&1; echo "chain exit $?" >> gpurun_out/chain_debug3.log
grep -n "NaN (row\|    row \|first non-finite\|chain exit" gpurun_out/chain_debug3.log | head -20
