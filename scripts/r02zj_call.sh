#!/bin/bash
# Round 2, call zj: dynamic hand-out of world batches to the persistent CTAs (default) against fixed shares (d0) on the
# bench's world count; warps per CTA at 8 192 worlds (the per-GPU share at N = 8); smoke().
set -u
mkdir -p gpurun_out
T=r02zj
timeout 600 python scripts/dev_sweep.py softbox 65536 200 k2:l8:d0 k2:l8 k2:l8:d0 k2:l8 > gpurun_out/${T}_sweep.log 2>&1
timeout 300 python scripts/dev_sweep.py softbox 8192 200 k2:l8:n16 k2:l8:n15 k2:l8:n14 k2:l8:n13 k2:l8 >> gpurun_out/${T}_sweep.log 2>&1
cat gpurun_out/${T}_sweep.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -n 2 gpurun_out/${T}_smoke.log
