#!/bin/bash
# Round 2, call g: trajectory kernels after the rewrite (SFU Box-Muller, two quads in flight, one-wave grids, parallel final
# reduction), their GPU parity tests, the regeneration with one launch per shape.
set -u
mkdir -p gpurun_out
T=r02g
python -m pytest tests/test_traj.py tests/test_gpu.py -m gpu -q -k "traj or regenerat or noise or stats or mask" > gpurun_out/${T}_pytest_traj.log 2>&1
python scripts/dev_traj_bench.py 65536 200 10 > gpurun_out/${T}_traj_bench.jsonl 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sg_traj -c 40 --csv \
    --log-file gpurun_out/${T}_traj_launches.csv python scripts/dev_traj_bench.py 65536 200 1 > /dev/null 2>&1
rm -rf /tmp/ds && ( time python soft-grip_b200/regenerate.py --out /tmp/ds --train 4096 --val 512 --test 512 \
   --softbox tests/golden/softbox.sgm --softball tests/golden/softball.sgm --softcylinder tests/golden/softcylinder.sgm \
   --tendon-damping softball=50 softcylinder=50 --noise-seed 3 ) > gpurun_out/${T}_regenerate.log 2>&1
tail -n 5 gpurun_out/${T}_pytest_traj.log
cat gpurun_out/${T}_traj_bench.jsonl gpurun_out/${T}_regenerate.log
