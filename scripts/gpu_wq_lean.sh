#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "weight_quant or s8 or segments" 2>&1 | tail -15
for t in 128 256; do
  echo "== OQ_WQ_LEAN_THREADS=$t"; OQ_WQ_LEAN_THREADS=$t timeout 300 python tools/step_time.py 2>&1 | tail -9
done
echo "== OQ_WQ_LEAN=0"; OQ_WQ_LEAN=0 timeout 300 python tools/step_time.py 2>&1 | tail -9
