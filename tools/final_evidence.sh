#!/bin/bash
# Round-end evidence on one B200: parity tests, smoke, the default bench line and the reference arm, ncu launch
# lists (bench command, one whole proof, the 2^20 MSM) and full captures of the dominant kernels.
# Everything lands in gpurun_out/; tools/kernel_shares.py and tools/ncu_summary.py turn it into profiles/.
set -u
T=${1:-fin}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${T}_gpu.txt 2>&1
lscpu | grep -E 'Model name|^CPU\(s\)' >> gpurun_out/${T}_gpu.txt
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/${T}_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== microbench"; timeout 120 tools/microbench > gpurun_out/${T}_microbench.json 2>&1; tail -c 400 gpurun_out/${T}_microbench.json
echo "== bench (default)"; timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; tail -c 600 gpurun_out/${T}_bench_default.json
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 600 gpurun_out/${T}_bench_reference.json
echo "== ncu launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${T}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_proof.csv \
    python tools/prove_bench.py 16 2 > gpurun_out/${T}_ncu_proof.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_msm20.csv \
    python tools/msm_sweep.py 20 0 > gpurun_out/${T}_ncu_msm20.log 2>&1
echo "== ncu full captures"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 4 -c 1 -f -o gpurun_out/${T}_prof_msm_accumulate_2_20 \
    python tools/msm_sweep.py 20 0 > gpurun_out/${T}_ncu_full_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 4 -c 2 -f -o gpurun_out/${T}_prof_ntt_2_20 \
    python tools/ntt_bench.py 20 1 > gpurun_out/${T}_ncu_full_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_quotient -s 3 -c 1 -f -o gpurun_out/${T}_prof_k_quotient \
    python tools/prove_bench.py 16 2 > gpurun_out/${T}_ncu_full_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_reduce_kernel -s 8 -c 1 -f -o gpurun_out/${T}_prof_msm_reduce \
    python tools/prove_bench.py 16 2 > gpurun_out/${T}_ncu_full_d.log 2>&1
ls -la gpurun_out | grep "${T}_"
