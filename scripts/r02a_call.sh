#!/bin/bash
# Round 2, first GPU call: baseline of everything (tests, bench on BASELINE configs[2] as written, configs[1], the
# trajectory kernels, the stabilised larger models with their lanes-per-world sweep, phase clocks).
set -u
mkdir -p gpurun_out
T=r02a
rm -f gpurun_out/test_gpu_measured.txt
python -c "import __graft_entry__ as g; g.smoke()"                 > gpurun_out/${T}_smoke.log 2>&1
python -m pytest tests -m gpu -q                                   > gpurun_out/${T}_pytest_gpu.log 2>&1
python bench.py                                                    > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --config 1 --steps 1                               > gpurun_out/${T}_bench_config1.json 2>> gpurun_out/${T}_bench.err
python scripts/dev_traj_bench.py 65536 200 10                      > gpurun_out/${T}_traj_bench.jsonl 2>&1
python scripts/dev_phase.py softbox 18944 l8:n16                   > gpurun_out/${T}_phase.log 2>&1
for L in 8 16; do
  SOFTGRIP_LPW=$L python bench.py --model softball --tendon-damping 50 --worlds 9472 --steps 1 --no-cpu-baseline --no-variants      > gpurun_out/${T}_bench_softball_lpw$L.json 2>> gpurun_out/${T}_bench.err
  SOFTGRIP_LPW=$L python bench.py --model softcylinder --tendon-damping 50 --worlds 9472 --steps 1 --no-cpu-baseline --no-variants  > gpurun_out/${T}_bench_softcylinder_lpw$L.json 2>> gpurun_out/${T}_bench.err
done
for L in 8 16 32; do
  SOFTGRIP_LPW=$L python bench.py --config 4 --worlds 4736 --steps 1 --no-cpu-baseline --no-variants > gpurun_out/${T}_bench_refined_lpw$L.json 2>> gpurun_out/${T}_bench.err
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sg_traj -c 40 --csv \
    --log-file gpurun_out/${T}_traj_launches.csv python scripts/dev_traj_bench.py 65536 200 1 > gpurun_out/${T}_traj_ncu.log 2>&1
tail -n 3 gpurun_out/${T}_pytest_gpu.log gpurun_out/${T}_smoke.log
cat gpurun_out/${T}_bench.json | cut -c1-1500
cat gpurun_out/test_gpu_measured.txt
tail -5 gpurun_out/${T}_bench.err
