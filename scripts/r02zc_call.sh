#!/bin/bash
# Round 2, call zc: ring addressed through its pinned shared-window address, uniform slider mass in a register, vector
# scans of the contact tables (default) against the previous commit (prev) and with the M^-1 block loaded before the
# cost test (mv); GPU tests of the default.
set -u
mkdir -p gpurun_out
T=r02zc
P=$PWD/soft-grip_b200
for v in prev default mv prev default mv; do
  echo "== $v" >> gpurun_out/${T}_sweep.log
  if [ $v = default ]; then python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
  else SOFTGRIP_LIB=$P/libsoftgrip_$v.so python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1; fi
done
cat gpurun_out/${T}_sweep.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/${T}_tests.log
