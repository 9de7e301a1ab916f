#!/bin/bash
# Round 2, call r: trajectory kernels with cached occupancy, 8 CTAs/SM for the fp32 noise kernel, buffers preallocated in the timing.
set -u
mkdir -p gpurun_out
T=r02r
python -m pytest tests/test_traj.py -m gpu -q > gpurun_out/${T}_pytest_traj.log 2>&1
python scripts/dev_traj_bench.py 65536 200 10 > gpurun_out/${T}_traj_bench.jsonl 2>&1
tail -n 3 gpurun_out/${T}_pytest_traj.log; cat gpurun_out/${T}_traj_bench.jsonl
