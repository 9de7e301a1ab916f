#!/bin/bash
# Round 2, call f: GPU tests on the new default kernel, the bench line, identical-worlds sweep (cost of unequal worlds in a
# warp), full dataset-tree regeneration timed (BASELINE configs[3]), ncu --set full of the trajectory kernels.
set -u
mkdir -p gpurun_out
T=r02f
rm -f gpurun_out/test_gpu_measured.txt
python -m pytest tests -m gpu -q                                   > gpurun_out/${T}_pytest_gpu.log 2>&1
python bench.py                                                    > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python scripts/dev_sweep.py softbox 9472 200 k2:l8 k2:l8:u1        > gpurun_out/${T}_sweep.log 2>&1
rm -rf /tmp/ds && ( time python soft-grip_b200/regenerate.py --out /tmp/ds --train 4096 --val 512 --test 512 \
   --softbox tests/golden/softbox.sgm --softball tests/golden/softball.sgm --softcylinder tests/golden/softcylinder.sgm \
   --tendon-damping softball=50 softcylinder=50 --noise-seed 3 ) > gpurun_out/${T}_regenerate.log 2>&1
python - >> gpurun_out/${T}_regenerate.log 2>&1 <<'PY'
import importlib, sys, glob, numpy as np
sys.path.insert(0, '.')
ds = importlib.import_module("soft-grip_b200.dataset")
for f in sorted(glob.glob('/tmp/ds/*/*.pickle')):
    x, k = ds.read_pickle(f)      # what functions/utils.py:8-23 does with a file
    print(f, x.shape, x.dtype, k.shape, 'finite', bool(np.isfinite(x).all()), 'k range %.0f..%.0f' % (k.min(), k.max()))
PY
ncu --set full --clock-control none --import-source on -k regex:sg_traj -c 6 -o gpurun_out/${T}_traj_full python scripts/dev_traj_bench.py 65536 200 1 > gpurun_out/${T}_traj_ncu.log 2>&1
tail -n 4 gpurun_out/${T}_pytest_gpu.log
cut -c1-400 gpurun_out/${T}_bench.json
cat gpurun_out/${T}_sweep.log gpurun_out/${T}_regenerate.log
tail -3 gpurun_out/${T}_bench.err gpurun_out/${T}_traj_ncu.log
