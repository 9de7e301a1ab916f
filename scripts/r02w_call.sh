#!/bin/bash
# Round 2, call w (8 GPUs): the bench line of the final kernel at N = 8 and N = 1 (rank 0's GPU), for the record.
set -u
mkdir -p gpurun_out
T=r02w
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
$TR bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/${T}_bench_8gpu.json 2> gpurun_out/${T}_bench_8gpu.err
python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
grep -h '^{' gpurun_out/${T}_bench_8gpu.json | cut -c1-300
grep -h '^{' gpurun_out/${T}_bench_1gpu.json | cut -c1-300
