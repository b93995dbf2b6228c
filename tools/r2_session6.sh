#!/bin/bash
# CUDA-graph replay of the prover rounds: parity first, then eager (B200_GRAPHS=0) vs replay on one proof alone and through the pool.
set -u
T=${1:-r2h}
mkdir -p gpurun_out
lscpu | grep -E 'Model name|^CPU\(s\)' > gpurun_out/${T}_host.txt
echo "== graph parity"; timeout 600 python -m pytest tests/test_gpu_graphs.py -x -q --timeout 500 2>&1 | tail -15 | tee gpurun_out/${T}_pytest_graphs.log
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 | tee gpurun_out/${T}_pytest_gpu.log
for g in 0 1; do
  echo "== B200_GRAPHS=$g one proof alone" | tee -a gpurun_out/${T}_graphs_ab.log
  for lg in 12 13 14 16; do B200_GRAPHS=$g timeout 200 python tools/prove_bench.py $lg 12 1 2>&1 | tail -1 | cut -c1-260 | tee -a gpurun_out/${T}_graphs_ab.log; done
  for mc in 8 32; do for c in 8 16; do
    echo "== B200_GRAPHS=$g CUDA_DEVICE_MAX_CONNECTIONS=$mc concurrency=$c" | tee -a gpurun_out/${T}_graphs_ab.log
    B200_GRAPHS=$g B200_BENCH_SKIP_LEGS=valid_match_mpc_collaborative CUDA_DEVICE_MAX_CONNECTIONS=$mc timeout 300 python bench.py --extras-only --no-cpu-baseline --concurrency $c 2>/dev/null | tail -1 | python -c '
import json, sys
d = json.loads(sys.stdin.readline())
for k, v in d.get("real_statements", {}).items():
    if isinstance(v, dict): print("  %-40s %8.1f proofs/s  one alone %.2f ms  kernels/proof %s  submit host us/proof %.0f" % (k[:40], v["proofs_per_s_e2e"], v["ms_one_proof_in_flight"], v.get("launches_per_proof"), v.get("launch_host_us_per_proof") or 0))
    else: print(k, v)
b = d.get("private_match_bundle", {})
print("  bundle", b.get("bundles_per_s_e2e"), b.get("error"))
' | tee -a gpurun_out/${T}_graphs_ab.log
  done; done
  echo "== B200_GRAPHS=$g headline" | tee -a gpurun_out/${T}_graphs_ab.log
  B200_GRAPHS=$g timeout 600 python bench.py --no-cpu-baseline --no-real-statements --no-msm 2>gpurun_out/${T}_bench_g$g.err | tail -1 > gpurun_out/${T}_bench_g$g.json
  python -c '
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches_per_proof","latency_ms_one_proof_in_flight")}, d["e2e"]["value"], d.get("steady_state",{}).get("value"), d["roofline"]["frac"])
' gpurun_out/${T}_bench_g$g.json 2>&1 | tee -a gpurun_out/${T}_graphs_ab.log
done
