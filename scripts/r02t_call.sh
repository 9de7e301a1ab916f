#!/bin/bash
# Round 2, call t: contact record carries the normalised friction block (default) against the previous record layout without
# pinned loop invariants (h0, 1.247e7 in r02s; with them 1.257e7); GPU tests.
set -u
mkdir -p gpurun_out
T=r02t
P=$PWD/soft-grip_b200
echo "== variant h0 (previous record layout, SG_HOIST=0)" >> gpurun_out/${T}_sweep.log
SOFTGRIP_LIB=$P/libsoftgrip_h0.so python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
echo "== default" >> gpurun_out/${T}_sweep.log
python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1
cat gpurun_out/${T}_sweep.log | cut -c1-220; tail -n 3 gpurun_out/${T}_pytest_gpu.log
