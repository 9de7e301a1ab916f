#!/bin/bash
# Round 2, call zi: evidence of the final kernel on one GPU -- GPU tests, the bench line, the ncu launch list of the bench
# command, DRAM bytes of one bench launch, one ncu --set full capture over a whole short episode.
set -u
mkdir -p gpurun_out
T=r02zi
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/${T}_tests.log
python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; grep -h '^{' gpurun_out/${T}_bench_1gpu.json | cut -c1-300
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; grep -h '^{' gpurun_out/${T}_bench_ref.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/${T}_bench_under_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sg_step_kernel2 -s 1 -c 1 --csv \
    --log-file gpurun_out/${T}_dram.csv python bench.py --steps 1 --warmup 3 --no-variants --no-cpu-baseline > /dev/null 2>&1
PROF_SETTLE=40 PROF_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:sg_step_kernel2 -c 1 \
   -o gpurun_out/${T}_k2_full python scripts/dev_prof.py softbox 9472 100 > gpurun_out/${T}_ncu.log 2>&1
tail -n 4 gpurun_out/${T}_dram.csv | cut -c1-300; tail -n 2 gpurun_out/${T}_ncu.log
