#!/bin/bash
# Round 2, call zh: the warps of a CTA step together in 1 / 2 / 4 / 8 groups (named barriers; SOFTGRIP_STEP_BARRIER).
set -u
mkdir -p gpurun_out
T=r02zh
timeout 900 python scripts/dev_sweep.py softbox 9472 200 k2:l8:s1 k2:l8:s2 k2:l8:s4 k2:l8:s8 k2:l8:s1 k2:l8:s2 k2:l8:s4 > gpurun_out/${T}_sweep.log 2>&1
cat gpurun_out/${T}_sweep.log | cut -c1-200
