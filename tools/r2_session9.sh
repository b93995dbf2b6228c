#!/bin/bash
# Host-side Horner over the bit sums + block-level batch inversion: parity (whole suite), one proof alone, extras; the pool tests
# (several threads capturing and replaying graphs at once) under compute-sanitizer.
set -u
T=${1:-r2l}
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 | tee gpurun_out/${T}_pytest_gpu.log
for lg in 12 13 14 16; do timeout 200 python tools/prove_bench.py $lg 12 1 2>&1 | tail -1 | cut -c1-260 | tee -a gpurun_out/${T}_prove_bench.log; done
timeout 200 python tools/msm_sweep.py 12,13,14,16,18,20 1 2>&1 | tee gpurun_out/${T}_msm_sweep_latency_plan.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c '
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches_per_proof","graph_launches_per_proof","latency_ms_one_proof_in_flight")}, d["e2e"]["value"])
for k,v in d["real_statements"].items(): print(k[:30], v["proofs_per_s_e2e"], v["ms_one_proof_in_flight"])
print("bundle", d["private_match_bundle"]["bundles_per_s_e2e"], d["private_match_bundle"]["cpu_baseline"])
' gpurun_out/${T}_bench.json
echo "== memcheck (pool tests: threads capturing / replaying graphs concurrently)"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pool.py -q --timeout 1400 > gpurun_out/${T}_memcheck_pool.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${T}_memcheck_pool.log | head -4
