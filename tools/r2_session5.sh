#!/bin/bash
# Round-2 session 5 (1 GPU): whole parity suite without -x, racecheck of a small proof, concurrency sweep.
set -u
T=${1:-r2e}
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -12 | tee gpurun_out/${T}_pytest_gpu.log
echo "== racecheck (2^12 proof)"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/prove_bench.py 12 1 > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard|Error" gpurun_out/${T}_racecheck.log | head -8
echo "== memcheck (2^12 proof)"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/prove_bench.py 12 1 > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY" gpurun_out/${T}_memcheck.log | head -3
for c in 4 6 8 10 12; do
  timeout 300 python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-msm --no-real-statements --concurrency $c 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'concurrency': d['run']['concurrency_per_gpu'], 'proofs_per_s': round(d['value'], 1), 'e2e': round(d['e2e']['value'], 1), 'e2e_pageable': round(d['e2e_pageable']['value'], 1), 'steady': d['steady_state'] and round(d['steady_state']['value'], 1), 'launches_per_proof': d.get('gpu_launches_per_proof')}))" | tee -a gpurun_out/${T}_concurrency_sweep.log
done
