#!/bin/bash
# Round 2, call n: latency-ordered friction solve (SG_FRIC_V2) and speculative block update + tree residual (SG_BLOCK_SPEC):
# p0 = both off (previous kernel), p1 = friction only, p2 = block only, default = both.
set -u
mkdir -p gpurun_out
T=r02n
P=$PWD/soft-grip_b200
for v in p0 p1 p2; do
  echo "== variant $v" >> gpurun_out/${T}_sweep.log
  SOFTGRIP_LIB=$P/libsoftgrip_$v.so python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
done
echo "== default (both)" >> gpurun_out/${T}_sweep.log
python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.log 2>&1
cat gpurun_out/${T}_sweep.log | cut -c1-200; tail -n 3 gpurun_out/${T}_pytest_gpu.log
