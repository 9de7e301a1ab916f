#!/bin/bash
# Round 2, call u: the final kernel (after the chain-loop work of r02s-t) -- smoke, GPU tests, the full bench line, phase clocks.
set -u
mkdir -p gpurun_out
T=r02u
rm -f gpurun_out/test_gpu_measured.txt
python -c "import __graft_entry__ as g; g.smoke()"                 > gpurun_out/${T}_smoke.log 2>&1
python -m pytest tests -m gpu -q                                   > gpurun_out/${T}_pytest_gpu.log 2>&1
python bench.py                                                    > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python scripts/dev_phase.py softbox 9472 l8:n16                    > gpurun_out/${T}_phase.log 2>&1
python scripts/dev_sweep.py softbox 9472 200 k2:l8                 > gpurun_out/${T}_sweep.log 2>&1
tail -n 3 gpurun_out/${T}_smoke.log gpurun_out/${T}_pytest_gpu.log
cut -c1-250 gpurun_out/${T}_bench.json; cat gpurun_out/${T}_sweep.log; cut -c1-250 gpurun_out/${T}_phase.log
