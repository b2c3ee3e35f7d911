#!/bin/bash
timeout 1200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "calibration_variant" 2>&1 | grep -E "^E|Error|assert|passed|failed" | head -40
