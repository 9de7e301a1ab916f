#!/bin/bash
# Round 2, call v: base pointers of the once-per-step loops (rows, warm start, Euler, capsule centres) pinned in registers
# (default) against the previous commit (pv).
set -u
mkdir -p gpurun_out
T=r02v
P=$PWD/soft-grip_b200
echo "== variant pv (previous commit)" >> gpurun_out/${T}_sweep.log
SOFTGRIP_LIB=$P/libsoftgrip_pv.so python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
echo "== default" >> gpurun_out/${T}_sweep.log
python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
python scripts/dev_phase.py softbox 9472 l8:n16 > gpurun_out/${T}_phase.log 2>&1
cat gpurun_out/${T}_sweep.log gpurun_out/${T}_phase.log | cut -c1-250
