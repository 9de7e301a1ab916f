#!/bin/bash
# Round 2, call zd: once-per-step loops with their global loads batched (batched(), 8 elements per round) against the
# previous commit (prev): contact-free part (40 rows) and the whole episode; GPU tests of the default.
set -u
mkdir -p gpurun_out
T=r02zd
P=$PWD/soft-grip_b200
for v in prev default prev default; do
  echo "== $v" >> gpurun_out/${T}_sweep.log
  if [ $v = default ]; then L=""; else L=$P/libsoftgrip_$v.so; fi
  SOFTGRIP_LIB=$L python scripts/dev_sweep.py softbox 9472 40 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
  SOFTGRIP_LIB=$L python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
done
cat gpurun_out/${T}_sweep.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/${T}_tests.log
timeout 300 python scripts/dev_phase.py softbox 9472 l8:n16:t0 > gpurun_out/${T}_phase.txt 2>&1; tail -n 4 gpurun_out/${T}_phase.txt | cut -c1-400
