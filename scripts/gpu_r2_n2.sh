#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "n2 exit $?"; tail -3 gpurun_out/r2_bench_n2.err; cut -c1-600 gpurun_out/r2_bench_n2.json
