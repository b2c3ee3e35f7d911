#!/bin/bash
# Full GPU check: all -m gpu tests, smoke, default bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -15 gpurun_out/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 1500 python bench.py "$@" > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -5 gpurun_out/bench.err; cat gpurun_out/bench.log
