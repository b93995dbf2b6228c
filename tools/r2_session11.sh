#!/bin/bash
# Host-mailbox batch inversion: parity (whole suite), latencies, sanitizer on a proof, bench.
set -u
T=${1:-r2n}
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 | tee gpurun_out/${T}_pytest_gpu.log
for lg in 12 13 14 16; do timeout 200 python tools/prove_bench.py $lg 20 1 2>&1 | tail -1 | cut -c1-260 | tee -a gpurun_out/${T}_prove_bench.log; done
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/prove_bench.py 12 1 > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY" gpurun_out/${T}_memcheck.log | head -3
echo "== bench"; timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c '
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches_per_proof","graph_launches_per_proof","latency_ms_one_proof_in_flight")}, d["e2e"]["value"])
for k,v in d["real_statements"].items(): print(k[:30], v["proofs_per_s_e2e"], v["ms_one_proof_in_flight"])
print("bundle", d["private_match_bundle"]["bundles_per_s_e2e"], d["private_match_bundle"]["cpu_baseline"])
' gpurun_out/${T}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_proof13.csv \
    python tools/prove_bench.py 13 2 1 > gpurun_out/${T}_ncu_proof13.log 2>&1
