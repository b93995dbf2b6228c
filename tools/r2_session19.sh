#!/bin/bash
# After the __syncwarp before the in-place stores of xyzz_add_quad: racecheck again, parity of the MSM / prover / real-SRS suites.
set -u
T=${1:-r2w}
mkdir -p gpurun_out
{
  echo "## racecheck, one 2^12 proof (latency plan: quad additions in block tree + bit sums)"
  timeout 100 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/prove_bench.py 12 1 1 > gpurun_out/${T}_racecheck_proof.log 2>&1; echo "rc=$?"
  grep -E "RACECHECK SUMMARY|Race reported|Error" gpurun_out/${T}_racecheck_proof.log | sort | uniq -c | head -8
  echo "## racecheck, lone MSMs 2^12 / 2^13 (latency plan: msm_tree_quad_kernel as well)"
  timeout 100 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/msm_sweep.py 12,13 1 > gpurun_out/${T}_racecheck_msm.log 2>&1; echo "rc=$?"
  grep -E "RACECHECK SUMMARY|Race reported|Error|same_result" gpurun_out/${T}_racecheck_msm.log | cut -c1-160 | sort | uniq -c | head -8
} 2>&1 | tee gpurun_out/${T}_sanitizer_summary.txt
echo "== parity"; timeout 200 python -m pytest tests/test_gpu_msm.py tests/test_gpu_real_srs.py tests/test_gpu_plonk.py tests/test_gpu_pool.py -q -x --timeout 180 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_subset.log
head -c 3000 gpurun_out/${T}_racecheck_proof.log | grep -A12 "Race reported" | head -30
