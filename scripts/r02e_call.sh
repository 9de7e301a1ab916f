#!/bin/bash
# Round 2, call e: broadphase run culling (t2, warp-uniform decision), + branch-free contact block (t4); timing
# experiments x1 / x2 on top of t4 (friction Newton capped at 1 / 0 iterations: wrong results, latency attribution only).
set -u
mkdir -p gpurun_out
T=r02e
P=$PWD/soft-grip_b200
for v in v01 t2 t4 x1 x2; do
  echo "== variant $v" >> gpurun_out/${T}_sweep.log
  SOFTGRIP_LIB=$P/libsoftgrip_$v.so python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
done
for v in t4; do
  echo "== variant $v" >> gpurun_out/${T}_phase.log
  SOFTGRIP_LIB=$P/libsoftgrip_$v.so python scripts/dev_phase.py softbox 9472 l8:n16 >> gpurun_out/${T}_phase.log 2>&1
done
echo "== geometry at 8192 worlds (t4): chosen, forced 16" >> gpurun_out/${T}_sweep.log
SOFTGRIP_LIB=$P/libsoftgrip_t4.so python scripts/dev_sweep.py softbox 8192 200 k2:l8 k2:l8:n16 >> gpurun_out/${T}_sweep.log 2>&1
grep -v "^\[W\|^  File\|^    " gpurun_out/${T}_sweep.log gpurun_out/${T}_phase.log | cut -c1-260
