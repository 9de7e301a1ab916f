#!/bin/bash
# Round 2, call zf: 4-byte slots in the tensor-memory sweep + slider state staged in shared memory for the row set-up
# (default) against the previous commit (prev) and with the staging off (g0); contact-free part and whole episode;
# GPU tests of the default.
set -u
mkdir -p gpurun_out
T=r02zf
P=$PWD/soft-grip_b200
for v in prev default g0 prev default g0; do
  echo "== $v" >> gpurun_out/${T}_sweep.log
  C=k2:l8; L=""
  if [ $v = prev ]; then L=$P/libsoftgrip_prev.so; fi
  if [ $v = g0 ]; then C=k2:l8:g0; fi
  SOFTGRIP_LIB=$L python scripts/dev_sweep.py softbox 9472 40 $C >> gpurun_out/${T}_sweep.log 2>&1
  SOFTGRIP_LIB=$L python scripts/dev_sweep.py softbox 9472 200 $C >> gpurun_out/${T}_sweep.log 2>&1
done
cat gpurun_out/${T}_sweep.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/${T}_tests.log
