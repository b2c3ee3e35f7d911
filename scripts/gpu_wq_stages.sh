#!/bin/bash
for st in 3 4 6; do echo "== stages $st"; OQ_WQ_WS_STAGES=$st timeout 300 python tools/wq_bench.py pc 2>&1 | tail -4; done
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "weight_quant or segments" 2>&1 | tail -3
