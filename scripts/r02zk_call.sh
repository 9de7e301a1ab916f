#!/bin/bash
# Round 2, call zk (2 GPUs): the bench line of the final kernel under torchrun at N = 2.
set -u
mkdir -p gpurun_out
T=r02zk
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
$TR bench.py --gpus 2 --steps 3 --warmup 3 --no-variants > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
grep -h '^{' gpurun_out/${T}_bench_2gpu.json | cut -c1-400; tail -n 3 gpurun_out/${T}_bench_2gpu.err
