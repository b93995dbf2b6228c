#!/bin/bash
# Round-2 session 1: toolchain probe, parity tests on the reworked MSM, per-size MSM / proof timings, pipe calibration.
set -u
mkdir -p gpurun_out
{
  echo "# toolchain probe on the GPU box ($(date -u +%FT%TZ))"
  for t in cargo rustc rustup go javac node; do printf '%s: ' $t; (command -v $t || echo "not found"); done
  ls -d /root/.cargo /usr/local/cargo /root/reference 2>&1
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv
  lscpu | grep -E 'Model name|^CPU\(s\)'
} > gpurun_out/r2_toolchain_probe.txt 2>&1
cat gpurun_out/r2_toolchain_probe.txt
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -15 | tee gpurun_out/r2a_pytest_gpu.log
echo "== microbench"; timeout 120 tools/microbench 2>&1 | tee gpurun_out/r2a_microbench.json
echo "== msm sweep"; timeout 600 python tools/msm_sweep.py 12,14,16,18,20 0 2>&1 | tee gpurun_out/r2a_msm_sweep.log
for lg in 12 13 14 16; do
  echo "== prove_bench $lg"; timeout 300 python tools/prove_bench.py $lg 10 2>&1 | tail -2 | tee -a gpurun_out/r2a_prove_bench.log
done
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
