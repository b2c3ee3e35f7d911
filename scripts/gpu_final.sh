#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1700 python bench.py --steps 4 --warmup 3 > gpurun_out/r2_bench_n1_v4.json 2> gpurun_out/r2_bench_n1_v4.err; echo "bench exit $?"; tail -2 gpurun_out/r2_bench_n1_v4.err; cut -c1-160 gpurun_out/r2_bench_n1_v4.json
