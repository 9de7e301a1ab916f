#!/bin/bash
# Round 2, call ze: ncu --set full of a contact-free launch (12 settle rows = 85 physics steps, 9 472 worlds): where the
# once-per-step phases of a contact-free step wait.
set -u
mkdir -p gpurun_out
T=r02ze
PROF_SETTLE=12 PROF_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sg_step_kernel2 -s 1 -c 1 \
   -o gpurun_out/${T}_k2_free python scripts/dev_prof.py softbox 9472 12 > gpurun_out/${T}_ncu.log 2>&1
tail -n 3 gpurun_out/${T}_ncu.log
