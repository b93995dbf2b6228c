#!/bin/bash
# Which batch inversion for the grand-product ratio?  host mailbox (default) vs one-lane binary Euclid on the device.
set -u
T=${1:-r2o}
mkdir -p gpurun_out
for inv in mailbox device; do
  echo "== B200_INVERSE=$inv" | tee -a gpurun_out/${T}_inverse_ab.log
  B200_INVERSE=$inv timeout 300 python -m pytest tests/test_gpu_field.py tests/test_gpu_plonk.py tests/test_gpu_link.py tests/test_gpu_graphs.py -q --timeout 280 2>&1 | tail -2 | tee -a gpurun_out/${T}_inverse_ab.log
  for lg in 12 13 14 16; do B200_INVERSE=$inv timeout 200 python tools/prove_bench.py $lg 20 1 2>&1 | tail -1 | cut -c1-250 | tee -a gpurun_out/${T}_inverse_ab.log; done
done
B200_INVERSE=device timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_proof13_device_inverse.csv \
    python tools/prove_bench.py 13 2 1 > gpurun_out/${T}_ncu_proof13.log 2>&1
echo "== pytest gpu (all, default)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4 | tee gpurun_out/${T}_pytest_gpu.log
