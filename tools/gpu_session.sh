#!/bin/bash
# One GPU-box session: parity tests, integer-pipe calibration, bench, ncu launch list + one full
# capture of the dominant kernel.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | grep -E 'Model name|^CPU\(s\)' >> gpurun_out/gpu.txt
echo "== pytest gpu" ; timeout 700 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== microbench" ; timeout 120 tools/microbench 2>&1 | tee gpurun_out/microbench.json
echo "== bench" ; timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.json
if [ "${1:-}" = "ncu" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "== ncu full capture of msm_accumulate_kernel"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 3 -c 1 -o gpurun_out/prof_msm_acc -f \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out
fi
