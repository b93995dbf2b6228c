#!/bin/bash
# Are the reference's real sizes (2^12 .. 2^14) GPU-bound or submission-bound in the pool?  The extras legs of bench.py
# (real statements, private-match bundle) at several pool widths and hardware-queue counts; the library reports the host
# time spent submitting launches (b200_launch_host_ns) beside the rates.
set -u
T=${1:-r2g}
mkdir -p gpurun_out
OUT=gpurun_out/${T}_small_proof_sweep.log
: > $OUT
for mc in 8 32; do
  for c in 8 16 32; do
    echo "== CUDA_DEVICE_MAX_CONNECTIONS=$mc concurrency=$c" | tee -a $OUT
    B200_BENCH_SKIP_LEGS=valid_match_mpc_collaborative CUDA_DEVICE_MAX_CONNECTIONS=$mc timeout 300 python bench.py --extras-only --no-cpu-baseline --concurrency $c 2>/dev/null | tail -1 | python -c '
import json, sys
d = json.loads(sys.stdin.readline())
for k, v in d.get("real_statements", {}).items():
    if isinstance(v, dict): print("  %-40s %8.1f proofs/s  one alone %.2f ms  launches/proof %s  launch host us/proof %s" % (k[:40], v["proofs_per_s_e2e"], v["ms_one_proof_in_flight"], v.get("launches_per_proof"), v.get("launch_host_us_per_proof")))
b = d.get("private_match_bundle", {})
print("  bundle", b.get("bundles_per_s_e2e"), b.get("error"))
' | tee -a $OUT
  done
done
