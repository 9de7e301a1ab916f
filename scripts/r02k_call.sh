#!/bin/bash
# Round 2, call k: lean broadphase chunk loop + uniform inverse weight in the row loop (default build) against the
# previous kernel with the same equality unroll (u4); phase clocks; GPU tests.
set -u
mkdir -p gpurun_out
T=r02k
P=$PWD/soft-grip_b200
echo "== variant u4 (previous kernel, equality sweep unrolled by 4)" >> gpurun_out/${T}_sweep.log
SOFTGRIP_LIB=$P/libsoftgrip_u4.so python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
echo "== default" >> gpurun_out/${T}_sweep.log
python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
python scripts/dev_phase.py softbox 9472 l8:n16 > gpurun_out/${T}_phase.log 2>&1
python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.log 2>&1
cat gpurun_out/${T}_sweep.log gpurun_out/${T}_phase.log | cut -c1-250; tail -n 3 gpurun_out/${T}_pytest_gpu.log
