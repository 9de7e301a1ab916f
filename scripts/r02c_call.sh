#!/bin/bash
# Round 2, call c: occupancy scan of the phase clocks (1..16 warps per CTA, one CTA per SM): per-warp-step cycles of every
# phase as a function of resident warps separates dependent-issue latency from issue-slot contention.
set -u
mkdir -p gpurun_out
T=r02c
P=$PWD/soft-grip_b200
for n in 1 2 4 8 12 16; do
  W=$((148 * 4 * n))
  SOFTGRIP_LIB=$P/libsoftgrip_v01.so python scripts/dev_phase.py softbox $W l8:n$n >> gpurun_out/${T}_phase_scan.log 2>&1
done
cat gpurun_out/${T}_phase_scan.log
