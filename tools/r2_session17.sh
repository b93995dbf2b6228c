#!/bin/bash
# xyzz_add_quad, second pass: the serial tree level takes the quad form only below kQuadTreeMaxOutputs; pool legs and headline per mode.
set -u
T=${1:-r2t}
mkdir -p gpurun_out
L=gpurun_out/${T}_quad_add_ab.log
echo "== parity, B200_QUAD_ADD=2 (hybrid)" | tee -a $L
B200_QUAD_ADD=2 timeout 420 python -m pytest tests/test_gpu_msm.py tests/test_gpu_plonk.py tests/test_gpu_graphs.py -q -x --timeout 400 2>&1 | tail -2 | tee -a $L
for m in 0 1 2; do
  echo "== one proof alone, B200_QUAD_ADD=$m" | tee -a $L
  for lg in 12 13 16; do B200_QUAD_ADD=$m timeout 200 python tools/prove_bench.py $lg 20 1 2>&1 | tail -1 | cut -c1-60 | tee -a $L; done
done
for m in 0 1 2 0 1 2; do
  echo "== pool legs (16 in flight), B200_QUAD_ADD=$m" | tee -a $L
  B200_QUAD_ADD=$m B200_BENCH_SKIP_LEGS=valid_match_mpc_collaborative timeout 300 python bench.py --extras-only --no-cpu-baseline --concurrency 16 2>/dev/null | tail -1 | python -c '
import json, sys
d = json.loads(sys.stdin.readline())
for k, v in d.get("real_statements", {}).items():
    if isinstance(v, dict): print("  %-40s %8.1f proofs/s  one alone %.2f ms" % (k[:40], v["proofs_per_s_e2e"], v["ms_one_proof_in_flight"]))
b = d.get("private_match_bundle", {})
print("  bundle", b.get("bundles_per_s_e2e"), b.get("error"))
' | tee -a $L
  echo "== headline, B200_QUAD_ADD=$m" | tee -a $L
  B200_QUAD_ADD=$m timeout 300 python bench.py --steps 100 --warmup 4 --no-cpu-baseline --no-msm --no-real-statements 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  headline', round(d['value'], 1), round(d['e2e']['value'], 1), 'one alone ms', round(d['latency_ms_one_proof_in_flight'], 3))" | tee -a $L
done
