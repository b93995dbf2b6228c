#!/bin/bash
# Round-2 session 4 (1 GPU): full parity suite (collaborative prover, block tree, CH = 16 scans), concurrency sweep, bench.
set -u
T=${1:-r2d}
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -15 | tee gpurun_out/${T}_pytest_gpu.log
echo "== msm sweep (throughput plan)"; timeout 300 python tools/msm_sweep.py 12,13,14,16,18,20 0 2>&1 | tee gpurun_out/${T}_msm_sweep_throughput_plan.log
echo "== msm sweep (latency plan)"; timeout 300 python tools/msm_sweep.py 12,13,14,16,18,20 1 2>&1 | tee gpurun_out/${T}_msm_sweep_latency_plan.log
echo "== ntt"; timeout 120 python tools/ntt_bench.py 16,20 1 2>&1 | tail -3 | tee gpurun_out/${T}_ntt_bench.log
for lg in 13 16; do
  echo "== prove_bench $lg (throughput plan / latency plan)"
  timeout 300 python tools/prove_bench.py $lg 10 0 2>&1 | tail -1 | tee -a gpurun_out/${T}_prove_bench.log
  timeout 300 python tools/prove_bench.py $lg 10 1 2>&1 | tail -1 | tee -a gpurun_out/${T}_prove_bench.log
done
for c in 4 6 8 10 12; do
  echo "== concurrency $c"; timeout 300 python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-msm --no-real-statements --concurrency $c 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'concurrency': d['run']['concurrency_per_gpu'], 'proofs_per_s': round(d['value'], 1), 'e2e': round(d['e2e']['value'], 1), 'launches_per_proof': d.get('gpu_launches_per_proof')}))" | tee -a gpurun_out/${T}_concurrency_sweep.log
done
echo "== bench"; timeout 1500 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json
d = json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print(json.dumps({k: d[k] for k in ('value', 'e2e', 'gpu_launches_per_proof', 'latency_ms_one_proof_in_flight')}))
print(json.dumps(d.get('real_statements'))[:1500]); print(json.dumps(d.get('private_match_bundle'))[:900])"; tail -3 gpurun_out/${T}_bench.err
