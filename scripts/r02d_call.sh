#!/bin/bash
# Round 2, call d: A/B of lean chain loop (t1), + broadphase run culling (t2), + unpredicated equality sweep (t3);
# timing experiments x1 / x2 (friction Newton capped at 1 / 0 iterations: wrong results, latency attribution only).
set -u
mkdir -p gpurun_out
T=r02d
P=$PWD/soft-grip_b200
for v in v01 t1 t2 t3 x1 x2; do
  echo "== variant $v" >> gpurun_out/${T}_sweep.log
  SOFTGRIP_LIB=$P/libsoftgrip_$v.so python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
done
for v in t2 t3; do
  echo "== variant $v" >> gpurun_out/${T}_phase.log
  SOFTGRIP_LIB=$P/libsoftgrip_$v.so python scripts/dev_phase.py softbox 9472 l8:n16 >> gpurun_out/${T}_phase.log 2>&1
done
echo "== geometry at 8192 worlds (t2)" >> gpurun_out/${T}_sweep.log
SOFTGRIP_LIB=$P/libsoftgrip_t2.so python scripts/dev_sweep.py softbox 8192 200 k2:l8 k2:l8:n16 >> gpurun_out/${T}_sweep.log 2>&1
cat gpurun_out/${T}_sweep.log gpurun_out/${T}_phase.log
