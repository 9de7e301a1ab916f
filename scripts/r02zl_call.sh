#!/bin/bash
# Round 2, call zl (8 GPUs): the bench line of the final kernel under torchrun at N = 8 (no CPU baseline, no variants: what is
# left of the round's GPU budget allows one short call).
set -u
mkdir -p gpurun_out
T=r02zl
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 55 $TR bench.py --gpus 8 --steps 3 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/${T}_bench_8gpu.json 2> gpurun_out/${T}_bench_8gpu.err
grep -h '^{' gpurun_out/${T}_bench_8gpu.json | cut -c1-300; tail -n 2 gpurun_out/${T}_bench_8gpu.err | cut -c1-200
