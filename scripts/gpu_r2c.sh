#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -15 gpurun_out/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 600 python tools/e2e_chain_debug.py 3 20 128 > gpurun_out/chain_debug.log 2>&1; echo "chain exit $?" >> gpurun_out/chain_debug.log
grep -n "oq debug\|FAILED\|param\|Wq\|qstat\|non-finite\|step\|exit" gpurun_out/chain_debug.log | head -60
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -5 gpurun_out/bench.err; cat gpurun_out/bench.log
