#!/bin/bash
# Round 2, call i (8 GPUs): the bench line at N = 8 (65 536 worlds sharded: 8 192 per GPU), the dataset tree over 8 ranks, the
# 1 M-world stress run of the refined composite.
set -u
mkdir -p gpurun_out
T=r02i
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/${T}_bench_8gpu.json 2> gpurun_out/${T}_bench_8gpu.err
rm -rf /tmp/ds8 && ( time $TR soft-grip_b200/regenerate.py --out /tmp/ds8 --train 4096 --val 512 --test 512 \
   --softbox tests/golden/softbox.sgm --softball tests/golden/softball.sgm --softcylinder tests/golden/softcylinder.sgm \
   --tendon-damping softball=50 softcylinder=50 --noise-seed 3 ) > gpurun_out/${T}_regenerate_8gpu.log 2>&1
ls -la /tmp/ds8/*/ >> gpurun_out/${T}_regenerate_8gpu.log 2>&1
timeout 600 $TR scripts/stress_1m.py 1048576 > gpurun_out/${T}_stress_1m.json 2> gpurun_out/${T}_stress_1m.err
grep -h '^{' gpurun_out/${T}_bench_8gpu.json | cut -c1-600
grep -v "^\[W\|Warning\|warn" gpurun_out/${T}_regenerate_8gpu.log | tail -25
cat gpurun_out/${T}_stress_1m.json; tail -3 gpurun_out/${T}_stress_1m.err gpurun_out/${T}_bench_8gpu.err
