#!/bin/bash
# compute-sanitizer over the graph-replayed prover and the evaluation-domain link proof; concurrency sweep with graphs on.
set -u
T=${1:-r2j}
mkdir -p gpurun_out
echo "== memcheck (2^12 proof x 8: eager, captured, replayed)"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/prove_bench.py 12 1 > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY" gpurun_out/${T}_memcheck.log | head -3
echo "== racecheck (same)"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/prove_bench.py 12 1 > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard|Error" gpurun_out/${T}_racecheck.log | head -8
echo "== memcheck (link proofs, eager and replayed)"; timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_link.py tests/test_gpu_graphs.py::test_link_proofs_replayed -q --timeout 1100 > gpurun_out/${T}_memcheck_link.log 2>&1; echo "memcheck link rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${T}_memcheck_link.log | head -4
{
  for f in memcheck racecheck memcheck_link; do echo "## ${T}_$f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/${T}_$f.log; done
} > gpurun_out/${T}_sanitizer_summary.txt
for c in 4 6 8 12; do
  timeout 300 python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-msm --no-real-statements --concurrency $c 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'concurrency': d['run']['concurrency_per_gpu'], 'proofs_per_s': round(d['value'], 1), 'e2e': round(d['e2e']['value'], 1), 'e2e_pageable': round(d['e2e_pageable']['value'], 1), 'steady': d['steady_state'] and round(d['steady_state']['value'], 1), 'kernels_per_proof': d.get('gpu_launches_per_proof'), 'graph_launches_per_proof': d.get('graph_launches_per_proof')}))" | tee -a gpurun_out/${T}_concurrency_sweep.log
done
