#!/bin/bash
# First gpurun call of the next round: everything that was written after round 1's GPU budget was spent, measured once.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/round2_first_call.sh'
# Outputs land in gpurun_out/ (copy what should be judged into profiles/, named per round).
set -u
mkdir -p gpurun_out
T=r02a
python -c "import __graft_entry__ as g; g.smoke()"                                  > gpurun_out/${T}_smoke.log 2>&1
python -m pytest tests -m gpu -x -q                                                 > gpurun_out/${T}_pytest_gpu.log 2>&1
python bench.py                                                                     > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python scripts/dev_traj_bench.py 65536 200 10                                       > gpurun_out/${T}_traj_bench.jsonl 2>&1
# the stabilised models of SURVEY 8d cfg 3 and BASELINE configs[4]
python bench.py --model softball --tendon-damping 50 --worlds-per-gpu 9472 --no-cpu-baseline          > gpurun_out/${T}_bench_softball.json 2>> gpurun_out/${T}_bench.err
python bench.py --model softcylinder --tendon-damping 50 --worlds-per-gpu 9472 --no-cpu-baseline      > gpurun_out/${T}_bench_softcylinder.json 2>> gpurun_out/${T}_bench.err
python bench.py --model softbox_refined --tendon-damping 20 --worlds-per-gpu 4736 --no-cpu-baseline   > gpurun_out/${T}_bench_softbox_refined.json 2>> gpurun_out/${T}_bench.err
# lanes per world for the larger models: sweep steps drop from 113 to 73 (softball) and from 213 to 120 / 107
# (refined softbox) at 16 / 32 lanes
for L in 16 32; do
  SOFTGRIP_LPW=$L python bench.py --model softball --tendon-damping 50 --worlds-per-gpu 9472 --no-cpu-baseline        > gpurun_out/${T}_bench_softball_lpw$L.json 2>> gpurun_out/${T}_bench.err
  SOFTGRIP_LPW=$L python bench.py --model softbox_refined --tendon-damping 20 --worlds-per-gpu 4736 --no-cpu-baseline > gpurun_out/${T}_bench_softbox_refined_lpw$L.json 2>> gpurun_out/${T}_bench.err
done
# launch list of the trajectory kernels (per-launch times under ncu are cold-cache and serialised: shares only)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sg_traj -c 40 --csv \
    --log-file gpurun_out/${T}_traj_launches.csv python scripts/dev_traj_bench.py 65536 200 1 > gpurun_out/${T}_traj_ncu.log 2>&1
tail -n 3 gpurun_out/${T}_pytest_gpu.log gpurun_out/${T}_smoke.log
cat gpurun_out/${T}_bench.json gpurun_out/${T}_traj_bench.jsonl
