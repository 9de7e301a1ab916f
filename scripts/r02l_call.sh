#!/bin/bash
# Round 2, call l: two equality rows per lane and step (SG_EQ2, default build) + lean broadphase chunks against the previous
# kernel (u4); phase clocks; GPU tests; bench.
set -u
mkdir -p gpurun_out
T=r02l
P=$PWD/soft-grip_b200
echo "== variant u4 (previous kernel, one row per slot, equality sweep unrolled by 4)" >> gpurun_out/${T}_sweep.log
SOFTGRIP_LIB=$P/libsoftgrip_u4.so python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
echo "== default (SG_EQ2, SG_BP_LEAN)" >> gpurun_out/${T}_sweep.log
python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
python scripts/dev_phase.py softbox 9472 l8:n16 > gpurun_out/${T}_phase.log 2>&1
python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1
python bench.py --no-variants > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_sweep.log gpurun_out/${T}_phase.log | cut -c1-250; tail -n 3 gpurun_out/${T}_pytest_gpu.log; cut -c1-200 gpurun_out/${T}_bench.json
