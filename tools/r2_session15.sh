#!/bin/bash
# Final validation after the last clean-ups + hardware queue count for the pool legs.
set -u
T=${1:-r2r}
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4 | tee gpurun_out/${T}_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/${T}_smoke.log
for mc in 8 32 8 32; do
  echo "== CUDA_DEVICE_MAX_CONNECTIONS=$mc" | tee -a gpurun_out/${T}_max_connections.log
  B200_BENCH_SKIP_LEGS=valid_match_mpc_collaborative CUDA_DEVICE_MAX_CONNECTIONS=$mc timeout 300 python bench.py --extras-only --no-cpu-baseline --concurrency 16 2>/dev/null | tail -1 | python -c '
import json, sys
d = json.loads(sys.stdin.readline())
for k, v in d.get("real_statements", {}).items():
    if isinstance(v, dict): print("  %-40s %8.1f proofs/s" % (k[:40], v["proofs_per_s_e2e"]))
print("  bundle", d.get("private_match_bundle", {}).get("bundles_per_s_e2e"))
' | tee -a gpurun_out/${T}_max_connections.log
  CUDA_DEVICE_MAX_CONNECTIONS=$mc timeout 300 python bench.py --steps 100 --warmup 4 --no-cpu-baseline --no-msm --no-real-statements 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  headline', round(d['value'], 1), round(d['e2e']['value'], 1))" | tee -a gpurun_out/${T}_max_connections.log
done
