#!/bin/bash
# Round 2, call za: contact records through a per-lane shared-memory ring (cp.async from L2 one block ahead) -- GPU parity
# tests, then throughput against the same library with the ring off (r0) and with ring and tensor memory off (t0).
set -u
mkdir -p gpurun_out
T=r02za
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/${T}_tests.log
timeout 600 python scripts/dev_sweep.py softbox 9472 200 k2:l8:a0:t0 k2:l8:a0:r0 k2:l8:a0 k2:l8:a0:r0 k2:l8:a0 > gpurun_out/${T}_sweep.log 2>&1
cat gpurun_out/${T}_sweep.log
