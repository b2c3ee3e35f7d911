#!/bin/bash
# round 2, final kernels: headline config at N=8 (and N=2, N=4 on the same box)
mkdir -p gpurun_out
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 3 --warmup 2 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r2_bench_n${n}_v3.json 2> gpurun_out/r2_bench_n${n}_v3.err; echo "n$n exit $?"; tail -1 gpurun_out/r2_bench_n${n}_v3.err; cut -c1-200 gpurun_out/r2_bench_n${n}_v3.json
done
