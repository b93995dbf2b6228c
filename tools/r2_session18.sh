#!/bin/bash
# Latency plan: fan-in of the serial tree level (8 = default, 4, 2) now that the shared-memory tree levels are quad additions.
set -u
T=${1:-r2u}
mkdir -p gpurun_out
L=gpurun_out/${T}_tree_fanin_ab.log
for a in 3 2 1; do
  echo "== B200_TREE_FANIN_LOG=$a" | tee -a $L
  B200_TREE_FANIN_LOG=$a timeout 200 python -m pytest tests/test_gpu_msm.py -q -x --timeout 180 2>&1 | tail -1 | tee -a $L
  for lg in 12 13 14 16; do B200_TREE_FANIN_LOG=$a timeout 200 python tools/prove_bench.py $lg 20 1 2>&1 | tail -1 | cut -c1-60 | tee -a $L; done
  B200_TREE_FANIN_LOG=$a timeout 200 python tools/msm_sweep.py 12,13,14,16,18 1 2>&1 | tail -5 | cut -c1-200 | tee -a $L
done
