#!/bin/bash
# Round-2 session 2: full GPU parity suite, memcheck of one small proof, per-size timings, launch lists.
set -u
T=${1:-r2b}
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -15 | tee gpurun_out/${T}_pytest_gpu.log
echo "== memcheck (2^12 proof)"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/prove_bench.py 12 1 > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|Error" gpurun_out/${T}_memcheck.log | head -10
echo "== msm sweep (auto plan)"; timeout 600 python tools/msm_sweep.py 12,13,14,16,18,20 0 2>&1 | tee gpurun_out/${T}_msm_sweep.log
echo "== msm window sweep, small sizes"; timeout 600 python tools/msm_sweep.py 12,14,16 11,12,13,14,15,16,17 2>&1 | tee gpurun_out/${T}_msm_window_sweep_small.log
for lg in 12 13 14 16; do
  echo "== prove_bench $lg"; timeout 300 python tools/prove_bench.py $lg 10 2>&1 | tail -2 | tee -a gpurun_out/${T}_prove_bench.log
done
echo "== ncu launch list, 2^16 proof"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_proof16.csv \
    python tools/prove_bench.py 16 2 > gpurun_out/${T}_ncu_proof16.log 2>&1
echo "== ncu launch list, 2^13 proof"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_proof13.csv \
    python tools/prove_bench.py 13 2 > gpurun_out/${T}_ncu_proof13.log 2>&1
python tools/kernel_shares.py gpurun_out/${T}_launches_proof16.csv | head -45
python tools/kernel_shares.py gpurun_out/${T}_launches_proof13.csv | head -45
echo "== bench"; timeout 1200 python bench.py --steps 60 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 3000 gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
