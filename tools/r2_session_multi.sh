#!/bin/bash
# Round-2 multi-GPU session on N GPUs of one box (N = number visible): C-ABI multi-GPU parity tests, SURVEY config 5 sweep
# (strong scaling of a fixed-size MSM), the bench line under torchrun.  Every step has a tight timeout: a hang costs
# N x the wall time.
set -u
T=${1:-r2m}
GPUS=${2:-1,2}
NB=${3:-2}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/${T}_gpus.txt
echo "== pytest multi"; timeout 240 python -m pytest tests/test_gpu_multi.py -q -x --timeout 200 2>&1 | tail -8 | tee gpurun_out/${T}_pytest_multi.log
echo "== sweep"; timeout 420 python tools/msm_sweep_multi.py --sizes ${4:-12,14,16,18,20,22,24} --gpus $GPUS --warmup 5 --iters 20 --cpu-max ${5:-22} > gpurun_out/${T}_msm_sweep.jsonl 2> gpurun_out/${T}_msm_sweep.err; tail -2 gpurun_out/${T}_msm_sweep.jsonl | cut -c1-1800; tail -3 gpurun_out/${T}_msm_sweep.err
echo "== bench N=$NB"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NB --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NB --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_${NB}gpu.json 2> gpurun_out/${T}_bench_${NB}gpu.err
tail -c 2500 gpurun_out/${T}_bench_${NB}gpu.json | cut -c1-2500; tail -3 gpurun_out/${T}_bench_${NB}gpu.err
