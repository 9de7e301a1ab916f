#!/bin/bash
# Round 2, call zb: tensor-memory rows + record ring -- the new GPU test, the bench line, phase shares, and the contact-only
# ncu capture for comparison with r02y.
set -u
mkdir -p gpurun_out
T=r02zb
timeout 600 python -m pytest tests -m gpu -x -q -k "tensor_memory" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/${T}_tests.log
python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; grep -h '^{' gpurun_out/${T}_bench_1gpu.json | cut -c1-400
timeout 300 python scripts/dev_phase.py softbox 9472 l8:n16:t0 > gpurun_out/${T}_phase.txt 2>&1; tail -n 30 gpurun_out/${T}_phase.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sg_step_kernel2 -s 1 -c 1 \
   -o gpurun_out/${T}_k2_contact python scripts/dev_prof_contact.py 9472 21 > gpurun_out/${T}_ncu.log 2>&1
tail -n 2 gpurun_out/${T}_ncu.log
