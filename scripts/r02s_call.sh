#!/bin/bash
# Round 2, call s: loop invariants of the chain loop pinned in registers (default) against rebuilt from the parameter bank (h0).
set -u
mkdir -p gpurun_out
T=r02s
P=$PWD/soft-grip_b200
echo "== variant h0 (SG_HOIST=0)" >> gpurun_out/${T}_sweep.log
SOFTGRIP_LIB=$P/libsoftgrip_h0.so python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
echo "== default (SG_HOIST=1)" >> gpurun_out/${T}_sweep.log
python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
cat gpurun_out/${T}_sweep.log | cut -c1-220
