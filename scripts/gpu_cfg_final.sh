#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name exit $?"; tail -1 gpurun_out/$name.err | cut -c1-200; python -c "
import json; d=json.load(open('gpurun_out/$name.json')); print(d['metric'], d['value'], 'e2e', (d.get('e2e') or {}).get('value'), 'launches/step', d['launches_per_optimizer_step'], 'step ms', d['optimizer_step_ms_eager_instrumented'])"; }
run r2_bench_cfg1_n1_v6 --config 1 --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline
run r2_bench_cfg5_n1_reduced_v6 --config 5 --epochs 2 --nsamples 8 --steps 2 --warmup 2 --no-cpu-baseline --no-gpu-eager-baseline --e2e-blocks 2
run r2_bench_cfg2_n1_reduced_v6 --config 2 --epochs 2 --nsamples 16 --steps 2 --warmup 2 --no-cpu-baseline --no-gpu-eager-baseline --e2e-blocks 2
