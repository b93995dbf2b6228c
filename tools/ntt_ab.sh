#!/bin/bash
# A/B of the NTT pass kernel: one tile per block (default) vs persistent cp.async double-buffered (B200_NTT_PERSISTENT=1).
# Parity first (the NTT test file with the variant on), then device times of 2^20 x 1, 2^19 x 7, 2^16 x 42 and a whole proof.
set -u
T=${1:-r2f}
mkdir -p gpurun_out
{
  echo "# NTT pass kernel A/B ($(date -u +%FT%TZ)): default = one tile per block, 3 resident blocks per SM; persistent = cp.async double buffer, 2 blocks per SM"
  echo "== parity with the persistent variant on"
  B200_NTT_PERSISTENT=1 timeout 600 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_plonk.py -q --timeout 500 2>&1 | tail -3
  for v in 0 1; do
    echo "== B200_NTT_PERSISTENT=$v"
    B200_NTT_PERSISTENT=$v timeout 200 python tools/ntt_bench.py 20 1 2>&1 | tail -1
    B200_NTT_PERSISTENT=$v timeout 200 python tools/ntt_bench.py 19 7 2>&1 | tail -1
    B200_NTT_PERSISTENT=$v timeout 200 python tools/ntt_bench.py 16 42 2>&1 | tail -1
    B200_NTT_PERSISTENT=$v timeout 200 python tools/prove_bench.py 16 10 2>&1 | tail -1 | cut -c1-330
  done
} 2>&1 | tee gpurun_out/${T}_ntt_persistent_ab.log
