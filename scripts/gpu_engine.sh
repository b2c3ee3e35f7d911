#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -q -s --timeout 300 -p no:cacheprovider "$@" > gpurun_out/engine.log 2>&1
echo "pytest exit $?" >> gpurun_out/engine.log
tail -60 gpurun_out/engine.log
