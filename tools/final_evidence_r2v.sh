#!/bin/bash
# Round-2 closing evidence on the final code (quad additions in the reduction tails): parity suite, smoke, one proof alone, MSM
# size sweeps (both plans), both bench arms, ncu launch lists (bench command with B200_GRAPHS=0 — see profiles/README r2k —, whole
# proofs at 2^16 / 2^13), full captures of the kernels that changed, sanitizers on a proof and on a lone MSM (the quad tree level).
set -u
T=${1:-r2v}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${T}_gpu.txt 2>&1
lscpu | grep -E 'Model name|^CPU\(s\)' >> gpurun_out/${T}_gpu.txt
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 | tee gpurun_out/${T}_pytest_gpu.log
if grep -qE "failed|error" gpurun_out/${T}_pytest_gpu.log; then echo "parity suite not green: stopping here"; exit 1; fi
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/${T}_smoke.log
for lg in 12 13 14 16; do timeout 200 python tools/prove_bench.py $lg 20 1 2>&1 | tail -1 | tee -a gpurun_out/${T}_prove_bench_latency_plan.log | cut -c1-200; done
echo "== msm sweep (throughput plan)"; timeout 300 python tools/msm_sweep.py 12,13,14,16,18,20,22,24 0 2>&1 | tee gpurun_out/${T}_msm_size_sweep_1gpu.log | cut -c1-230
echo "== msm sweep (latency plan)"; timeout 300 python tools/msm_sweep.py 12,13,14,16,18,20 1 2>&1 | tee gpurun_out/${T}_msm_sweep_latency_plan.log | cut -c1-230
echo "== bench (default)"; timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; tail -c 600 gpurun_out/${T}_bench_default.json; tail -3 gpurun_out/${T}_bench_default.err
echo "== bench --steps 20 --warmup 3 (as the driver runs it)"; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_driver_like.json 2>/dev/null; tail -c 300 gpurun_out/${T}_bench_driver_like.json
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 300 gpurun_out/${T}_bench_reference.json
echo "== ncu launch lists"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_proof16.csv \
    python tools/prove_bench.py 16 2 1 > gpurun_out/${T}_ncu_proof16.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_proof13.csv \
    python tools/prove_bench.py 13 2 1 > gpurun_out/${T}_ncu_proof13.log 2>&1
B200_GRAPHS=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${T}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-real-statements > gpurun_out/${T}_ncu_bench.log 2>&1
echo "== ncu full captures"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msm_bitsum -s 8 -c 1 -f -o gpurun_out/${T}_prof_msm_bitsum \
    python tools/prove_bench.py 13 2 1 > gpurun_out/${T}_ncu_full_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msm_blocktree -s 8 -c 1 -f -o gpurun_out/${T}_prof_msm_blocktree \
    python tools/prove_bench.py 13 2 1 > gpurun_out/${T}_ncu_full_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 6 -c 1 -f -o gpurun_out/${T}_prof_msm_accumulate_proof \
    python tools/prove_bench.py 16 2 > gpurun_out/${T}_ncu_full_e.log 2>&1
echo "== sanitizers"
{
  echo "## racecheck, one 2^12 proof (latency plan: quad additions in block tree + bit sums)"
  timeout 240 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/prove_bench.py 12 1 1 2>&1 | grep -E "RACECHECK SUMMARY|hazard|Error" | head -6
  echo "## racecheck, lone MSMs 2^12 / 2^13 (latency plan: msm_tree_quad_kernel as well)"
  timeout 240 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/msm_sweep.py 12,13 1 2>&1 | grep -E "RACECHECK SUMMARY|hazard|Error|same_result" | cut -c1-200 | head -6
  echo "## memcheck, one 2^12 proof"
  timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/prove_bench.py 12 1 1 2>&1 | grep -E "ERROR SUMMARY" | head -3
} > gpurun_out/${T}_sanitizer_summary.txt 2>&1
cat gpurun_out/${T}_sanitizer_summary.txt
ls gpurun_out | grep -c "${T}_"
