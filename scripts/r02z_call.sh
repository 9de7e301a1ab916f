#!/bin/bash
# Round 2, call z: the (u, n) rows of the equality sweep in tensor memory -- GPU parity tests, then throughput against the
# shared-memory rows (SOFTGRIP_TMEM=0), same library, alternating.
set -u
mkdir -p gpurun_out
T=r02z
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/${T}_tests.log
timeout 600 python scripts/dev_sweep.py softbox 9472 200 k2:l8:a0:t0 k2:l8:a0 k2:l8:a0:t0 k2:l8:a0 > gpurun_out/${T}_sweep.log 2>&1
cat gpurun_out/${T}_sweep.log
