#!/bin/bash
# Round 2, call q: ncu --set full of the rewritten trajectory kernels.
set -u
mkdir -p gpurun_out
T=r02q
ncu --set full --clock-control none --import-source on -k regex:sg_traj -c 6 -o gpurun_out/${T}_traj_full python scripts/dev_traj_bench.py 65536 200 1 > gpurun_out/${T}_traj_ncu.log 2>&1
tail -2 gpurun_out/${T}_traj_ncu.log
