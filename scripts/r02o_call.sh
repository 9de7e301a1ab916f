#!/bin/bash
# Round 2, call o: record prefetch distance 1 (default) against 2 (d2).
set -u
mkdir -p gpurun_out
T=r02o
P=$PWD/soft-grip_b200
echo "== variant d2 (two blocks ahead)" >> gpurun_out/${T}_sweep.log
SOFTGRIP_LIB=$P/libsoftgrip_d2.so python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
echo "== default (one block ahead)" >> gpurun_out/${T}_sweep.log
python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
python scripts/dev_phase.py softbox 9472 l8:n16 > gpurun_out/${T}_phase.log 2>&1
cat gpurun_out/${T}_sweep.log gpurun_out/${T}_phase.log | cut -c1-250
