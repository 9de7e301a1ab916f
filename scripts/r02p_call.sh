#!/bin/bash
# Round 2, call p: timing experiment (wrong physics): contact-block update without the M^-1 multiply and its four shared-memory loads.
set -u
mkdir -p gpurun_out
T=r02p
P=$PWD/soft-grip_b200
echo "== variant xd (diagonal update, timing only)" >> gpurun_out/${T}_sweep.log
SOFTGRIP_LIB=$P/libsoftgrip_xd.so python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
echo "== default" >> gpurun_out/${T}_sweep.log
python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
cat gpurun_out/${T}_sweep.log | cut -c1-220
