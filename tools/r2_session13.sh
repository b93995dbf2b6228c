#!/bin/bash
# Final-code parity + how many proofs in flight the pool wants now that submission is cheap (graphs).
set -u
T=${1:-r2p}
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4 | tee gpurun_out/${T}_pytest_gpu.log
for c in 12 16 24 32; do
  echo "== concurrency=$c" | tee -a gpurun_out/${T}_pool_width.log
  B200_BENCH_SKIP_LEGS=valid_match_mpc_collaborative timeout 300 python bench.py --extras-only --no-cpu-baseline --concurrency $c 2>/dev/null | tail -1 | python -c '
import json, sys
d = json.loads(sys.stdin.readline())
for k, v in d.get("real_statements", {}).items():
    if isinstance(v, dict): print("  %-40s %8.1f proofs/s  one alone %.2f ms  submit host us/proof %.0f" % (k[:40], v["proofs_per_s_e2e"], v["ms_one_proof_in_flight"], v.get("launch_host_us_per_proof") or 0))
b = d.get("private_match_bundle", {})
print("  bundle", b.get("bundles_per_s_e2e"), b.get("error"))
' | tee -a gpurun_out/${T}_pool_width.log
done
for c in 6 8; do
  timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-msm --no-real-statements --concurrency $c 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'concurrency': d['run']['concurrency_per_gpu'], 'proofs_per_s': round(d['value'], 1), 'e2e': round(d['e2e']['value'], 1), 'one_alone_ms': round(d['latency_ms_one_proof_in_flight'], 3), 'kernels_per_proof': d.get('gpu_launches_per_proof'), 'graph_launches_per_proof': d.get('graph_launches_per_proof')}))" | tee -a gpurun_out/${T}_pool_width.log
done
