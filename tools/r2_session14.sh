#!/bin/bash
# The bench as the driver runs it: short (20 steps / 3 warm-up) and default; both arms; the pool legs three times (spread).
set -u
T=${1:-r2q}
mkdir -p gpurun_out
for a in "--steps 20 --warmup 3" "--steps 60 --warmup 5" ""; do
  echo "== bench.py --gpus 1 $a" | tee -a gpurun_out/${T}_bench_runs.log
  timeout 900 python bench.py --gpus 1 $a > gpurun_out/${T}_bench_tmp.json 2> gpurun_out/${T}_bench_tmp.err; python -c '
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(json.dumps({"value":round(d["value"],1),"e2e":round(d["e2e"]["value"],1),"steps":d["steps"],"warmup":d["warmup"],"ms_per_step":round(d["ms_per_step"],3),"steady":d.get("steady_state") and round(d["steady_state"]["value"],1),"kernels_per_proof":d["gpu_launches_per_proof"],"graphs_per_proof":d["graph_launches_per_proof"],"roofline_frac":round(d["roofline"]["frac"],5),
 "vbc":round(d["real_statements"]["valid_balance_create (BASELINE.json configs[0])"]["proofs_per_s_e2e"],1),"settlement":round(d["real_statements"]["intent_and_balance_private_settlement (the statement of configs[3])"]["proofs_per_s_e2e"],1),"bundle":round(d["private_match_bundle"]["bundles_per_s_e2e"],1),"bundle_bit_exact":d["private_match_bundle"]["cpu_baseline"]["bit_exact_vs_gpu"]}))
' gpurun_out/${T}_bench_tmp.json 2>&1 | tee -a gpurun_out/${T}_bench_runs.log; tail -2 gpurun_out/${T}_bench_tmp.err
done
cp gpurun_out/${T}_bench_tmp.json gpurun_out/${T}_bench_default.json
echo "== reference arm" | tee -a gpurun_out/${T}_bench_runs.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400 | tee -a gpurun_out/${T}_bench_runs.log
