#!/bin/bash
# Round-2 session 3 (1 GPU): full parity suite incl. bundle / service / real-SRS tests, plan check, bench.
set -u
T=${1:-r2c}
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -15 | tee gpurun_out/${T}_pytest_gpu.log
echo "== msm sweep (auto plan)"; timeout 600 python tools/msm_sweep.py 12,13,14,16,18,20 0 2>&1 | tee gpurun_out/${T}_msm_sweep.log
echo "== msm window sweep 2^20"; timeout 600 python tools/msm_sweep.py 20 17,19,20 2>&1 | tee gpurun_out/${T}_msm_window_sweep_20.log
for lg in 13 16; do
  echo "== prove_bench $lg"; timeout 300 python tools/prove_bench.py $lg 10 2>&1 | tail -2 | tee -a gpurun_out/${T}_prove_bench.log
done
echo "== bench"; timeout 1500 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 6000 gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 800 gpurun_out/${T}_bench_reference.json
