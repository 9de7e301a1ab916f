#!/bin/bash
# Round 2, call b: A/B of the chain-phase / step-slot variants (csrc/Makefile `variant`), lanes per world, and one
# ncu --set full capture of a contact-heavy window.
set -u
mkdir -p gpurun_out
T=r02b
P=$PWD/soft-grip_b200
for v in v00 v10 v01; do
  echo "== variant $v" >> gpurun_out/${T}_sweep.log
  SOFTGRIP_LIB=$P/libsoftgrip_$v.so python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
done
echo "== default build (v11)" >> gpurun_out/${T}_sweep.log
python scripts/dev_sweep.py softbox 9472 200 k2:l8 k2:l4 k2:l4:n8 k2:l16 >> gpurun_out/${T}_sweep.log 2>&1
python scripts/dev_phase.py softbox 9472 l8:n16 l4:n8 > gpurun_out/${T}_phase.log 2>&1
# contact-heavy ncu window: 100 rows (rows 40..100 carry 20-60 contacts)
PROF_SETTLE=40 PROF_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:sg_step_kernel2 -c 1 \
   -o gpurun_out/${T}_k2_full python scripts/dev_prof.py softbox 9472 100 > gpurun_out/${T}_ncu.log 2>&1
cat gpurun_out/${T}_sweep.log
cat gpurun_out/${T}_phase.log
tail -3 gpurun_out/${T}_ncu.log
