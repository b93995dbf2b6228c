#!/bin/bash
# Round-end evidence on one B200: parity tests, smoke, the default bench line and the reference arm, ncu launch lists (bench
# command, one whole proof at 2^16 and 2^13, the 2^20 MSM) and full captures of the dominant kernels.  Everything lands in
# gpurun_out/; tools/kernel_shares.py, tools/ncu_summary.py and tools/ncu_traffic.py turn it into profiles/.
set -u
T=${1:-fin}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${T}_gpu.txt 2>&1
lscpu | grep -E 'Model name|^CPU\(s\)' >> gpurun_out/${T}_gpu.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 | tee gpurun_out/${T}_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== microbench"; timeout 120 tools/microbench > gpurun_out/${T}_microbench.json 2>&1; tail -c 500 gpurun_out/${T}_microbench.json
echo "== msm sweep"; timeout 300 python tools/msm_sweep.py 12,13,14,16,18,20,22,24 0 2>&1 | tee gpurun_out/${T}_msm_size_sweep_1gpu.log
echo "== ntt"; timeout 120 python tools/ntt_bench.py 20 8 2>&1 | tail -3 | tee gpurun_out/${T}_ntt_bench.log
for lg in 12 13 14 16; do timeout 200 python tools/prove_bench.py $lg 10 1 2>&1 | tail -1 | tee -a gpurun_out/${T}_prove_bench_latency_plan.log; done
echo "== bench (default)"; timeout 1200 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; tail -c 700 gpurun_out/${T}_bench_default.json; tail -3 gpurun_out/${T}_bench_default.err
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 400 gpurun_out/${T}_bench_reference.json
echo "== ncu launch lists"
# B200_GRAPHS=0: ncu dies (SIGSEGV, no message) when several host threads capture streams while it serialises kernels
# (profiles/README.md, r2k); the kernels are the same eager or replayed, and one worker with graphs on profiles fine
B200_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${T}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-real-statements > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_proof16.csv \
    python tools/prove_bench.py 16 2 > gpurun_out/${T}_ncu_proof16.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_proof13.csv \
    python tools/prove_bench.py 13 2 > gpurun_out/${T}_ncu_proof13.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_msm20.csv \
    python tools/msm_sweep.py 20 0 > gpurun_out/${T}_ncu_msm20.log 2>&1
echo "== ncu full captures"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 4 -c 1 -f -o gpurun_out/${T}_prof_msm_accumulate_2_20 \
    python tools/msm_sweep.py 20 0 > gpurun_out/${T}_ncu_full_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 6 -c 1 -f -o gpurun_out/${T}_prof_msm_accumulate_proof \
    python tools/prove_bench.py 16 2 > gpurun_out/${T}_ncu_full_e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 4 -c 2 -f -o gpurun_out/${T}_prof_ntt_2_20 \
    python tools/ntt_bench.py 20 1 > gpurun_out/${T}_ncu_full_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_tree_kernel -s 8 -c 1 -f -o gpurun_out/${T}_prof_msm_tree \
    python tools/prove_bench.py 16 2 > gpurun_out/${T}_ncu_full_d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_blocktree_kernel -s 8 -c 1 -f -o gpurun_out/${T}_prof_msm_blocktree \
    python tools/prove_bench.py 16 2 > gpurun_out/${T}_ncu_full_f.log 2>&1
ls -la gpurun_out | grep "${T}_" | wc -l
