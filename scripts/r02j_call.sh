#!/bin/bash
# Round 2, call j: equality-sweep unroll A/B, then the evidence of the final kernel: ncu launch list of the bench command,
# one ncu --set full capture (contact-heavy window), DRAM bytes of one bench launch.
set -u
mkdir -p gpurun_out
T=r02j
P=$PWD/soft-grip_b200
for v in u1 u4; do
  echo "== variant $v (SG_EQ_UNROLL)" >> gpurun_out/${T}_sweep.log
  SOFTGRIP_LIB=$P/libsoftgrip_$v.so python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
done
echo "== default" >> gpurun_out/${T}_sweep.log
python scripts/dev_sweep.py softbox 9472 200 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/${T}_bench_under_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sg_step_kernel2 -s 1 -c 1 --csv \
    --log-file gpurun_out/${T}_dram.csv python bench.py --steps 1 --warmup 3 --no-variants --no-cpu-baseline > /dev/null 2>&1
PROF_SETTLE=40 PROF_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:sg_step_kernel2 -c 1 \
   -o gpurun_out/${T}_k2_full python scripts/dev_prof.py softbox 9472 100 > gpurun_out/${T}_ncu.log 2>&1
cat gpurun_out/${T}_sweep.log; cat gpurun_out/${T}_dram.csv | tail -4; tail -n 2 gpurun_out/${T}_ncu.log
