#!/bin/bash
# round 2: N=8 lines: the headline config, then configs 4 and 5 with their full epochs x samples
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name exit $?"; tail -2 gpurun_out/$name.err; cut -c1-300 gpurun_out/$name.json; }
run r2_bench_n8 --steps 3 --warmup 2
run r2_bench_cfg4_n8 --config 4 --steps 2 --warmup 1 --e2e-blocks 2
run r2_bench_cfg5_n8 --config 5 --steps 2 --warmup 1 --e2e-blocks 2
