#!/bin/bash
# ncu launch list of whole proofs + full captures of the top kernels (1 GPU; numbers printed under ncu are not bench values)
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_proof.csv \
    python tools/prove_bench.py 16 2 > gpurun_out/ncu_proof.log 2>&1
for k in ntt_pass_kernel k_quotient msm_accumulate msm_reduce_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 1 -f -o gpurun_out/prof_$k \
      python tools/prove_bench.py 16 2 > gpurun_out/ncu_full_$k.log 2>&1
done
ls -la gpurun_out | tail -8
