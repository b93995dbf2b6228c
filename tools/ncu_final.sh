#!/bin/bash
# Round-end evidence: launch list of whole proofs and of the 2^20 MSM, full captures of the dominant kernels.
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_proof_final.csv \
    python tools/prove_bench.py 16 2 > gpurun_out/ncu_proof_final.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_msm20_final.csv \
    python tools/msm_sweep.py 20 0 > gpurun_out/ncu_msm20_final.log 2>&1
# msm_accumulate of the 2^20 MSM (skip the warm-up launches), and the batched 7 x 2^19 coset NTT pass inside a proof
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 4 -c 1 -f -o gpurun_out/prof_final_msm_accumulate_2_20 \
    python tools/msm_sweep.py 20 0 > gpurun_out/ncu_full_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 4 -c 1 -f -o gpurun_out/prof_final_ntt_2_20 \
    python tools/ntt_bench.py 20 1 > gpurun_out/ncu_full_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_quotient -s 3 -c 1 -f -o gpurun_out/prof_final_k_quotient \
    python tools/prove_bench.py 16 2 > gpurun_out/ncu_full_c.log 2>&1
ls -la gpurun_out | grep -E "final"
