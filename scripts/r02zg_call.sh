#!/bin/bash
# Round 2, call zg: default (8-byte slots, staged slider state) against the previous commit (prev) and the 4-byte-slot
# build (tm4); whole episode; then the bench line and the GPU tests of the default.
set -u
mkdir -p gpurun_out
T=r02zg
P=$PWD/soft-grip_b200
for v in prev default tm4 prev default tm4; do
  echo "== $v" >> gpurun_out/${T}_sweep.log
  L=""
  if [ $v != default ]; then L=$P/libsoftgrip_$v.so; fi
  SOFTGRIP_LIB=$L python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
done
cat gpurun_out/${T}_sweep.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/${T}_tests.log
python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; grep -h '^{' gpurun_out/${T}_bench_1gpu.json | cut -c1-600
