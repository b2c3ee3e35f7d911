#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r2_bench_n2_v5.json 2> gpurun_out/r2_bench_n2_v5.err; echo "n2 exit $?"; tail -2 gpurun_out/r2_bench_n2_v5.err; cut -c1-200 gpurun_out/r2_bench_n2_v5.json
timeout 300 python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 2>/dev/null | cut -c1-300
