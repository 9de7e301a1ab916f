#!/bin/bash
# Round 2, call y: ncu --set full of a contact-only launch (every world at the 58-contact snapshot, 21 physics steps through the
# step API): the stall samples of the contact-rich steps, which the rollout captures miss (their sampling buffer fills in the
# contact-free settle phase).
set -u
mkdir -p gpurun_out
T=r02y
python scripts/dev_prof_contact.py 9472 21 > gpurun_out/${T}_plain.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sg_step_kernel2 -s 1 -c 1 \
   -o gpurun_out/${T}_k2_contact python scripts/dev_prof_contact.py 9472 21 > gpurun_out/${T}_ncu.log 2>&1
cat gpurun_out/${T}_plain.log; tail -n 2 gpurun_out/${T}_ncu.log
