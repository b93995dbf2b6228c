#!/bin/bash
# Where does a lone small proof spend its time now?  A/B of the host-side Horner, ncu launch list of a 2^13 proof.
set -u
T=${1:-r2m}
mkdir -p gpurun_out
for h in 1 0; do
  echo "== B200_HOST_HORNER=$h" | tee -a gpurun_out/${T}_host_horner_ab.log
  for lg in 12 13 16; do B200_HOST_HORNER=$h timeout 200 python tools/prove_bench.py $lg 20 1 2>&1 | tail -1 | cut -c1-250 | tee -a gpurun_out/${T}_host_horner_ab.log; done
  B200_HOST_HORNER=$h timeout 200 python tools/msm_sweep.py 13,16 1 2>&1 | tee -a gpurun_out/${T}_host_horner_ab.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_proof13.csv \
    python tools/prove_bench.py 13 2 1 > gpurun_out/${T}_ncu_proof13.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_proof16.csv \
    python tools/prove_bench.py 16 2 > gpurun_out/${T}_ncu_proof16.log 2>&1
