#!/bin/bash
# Round-2 multi-GPU session (one 8-GPU box): C-ABI multi-GPU parity tests, SURVEY config 5 sweep (strong scaling of a
# fixed-size MSM), the bench line at N = 2 and N = 8 under torchrun.
set -u
T=${1:-r2m}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/${T}_gpus.txt
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -q -x --timeout 600 2>&1 | tail -8 | tee gpurun_out/${T}_pytest_multi.log
echo "== sweep"; timeout 1500 python tools/msm_sweep_multi.py --sizes 12,14,16,18,20,22,24 --gpus 1,2,4,8 > gpurun_out/${T}_msm_sweep.jsonl 2> gpurun_out/${T}_msm_sweep.err; tail -3 gpurun_out/${T}_msm_sweep.jsonl | cut -c1-1500; tail -3 gpurun_out/${T}_msm_sweep.err
for N in 2 8; do
  echo "== bench N=$N"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
  tail -c 1800 gpurun_out/${T}_bench_${N}gpu.json | cut -c1-1800; tail -3 gpurun_out/${T}_bench_${N}gpu.err
done
