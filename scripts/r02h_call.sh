#!/bin/bash
# Round 2, call h: smoke + GPU tests + bench after the warm-start change.
set -u
mkdir -p gpurun_out
T=r02h
rm -f gpurun_out/test_gpu_measured.txt
python -c "import __graft_entry__ as g; g.smoke()"                 > gpurun_out/${T}_smoke.log 2>&1
python -m pytest tests -m gpu -q                                   > gpurun_out/${T}_pytest_gpu.log 2>&1
python bench.py --no-variants                                      > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -n 3 gpurun_out/${T}_smoke.log gpurun_out/${T}_pytest_gpu.log
cut -c1-300 gpurun_out/${T}_bench.json
tail -n 3 gpurun_out/${T}_bench.err
