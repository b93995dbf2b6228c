#!/bin/bash
# The bench command under ncu died with "free(): invalid size" once graphs were on: is it ncu (multi-threaded capture / external
# event nodes) or the library?  glibc heap checking on the plain run, then ncu with graphs off / on.
set -u
T=${1:-r2k}
mkdir -p gpurun_out
echo "== plain bench, glibc heap checks on"; MALLOC_CHECK_=3 MALLOC_PERTURB_=165 timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-msm > gpurun_out/${T}_bench_malloc_check.json 2> gpurun_out/${T}_bench_malloc_check.err; echo "rc=$?"; tail -c 300 gpurun_out/${T}_bench_malloc_check.json; tail -3 gpurun_out/${T}_bench_malloc_check.err
echo "== ncu launch list, graphs off"; B200_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${T}_launches_bench_eager.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-real-statements > gpurun_out/${T}_ncu_bench_eager.log 2>&1; echo "rc=$?"; tail -c 200 gpurun_out/${T}_ncu_bench_eager.log; wc -l gpurun_out/${T}_launches_bench_eager.csv
echo "== ncu launch list, graphs on"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${T}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-real-statements > gpurun_out/${T}_ncu_bench.log 2>&1; echo "rc=$?"; tail -c 200 gpurun_out/${T}_ncu_bench.log; wc -l gpurun_out/${T}_launches_bench.csv
echo "== ncu launch list, graphs on, one worker"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${T}_launches_bench_c1.csv \
    python bench.py --steps 2 --warmup 3 --concurrency 1 --no-cpu-baseline --no-real-statements > gpurun_out/${T}_ncu_bench_c1.log 2>&1; echo "rc=$?"; tail -c 200 gpurun_out/${T}_ncu_bench_c1.log; wc -l gpurun_out/${T}_launches_bench_c1.csv
bash tools/r2_session7.sh ${T}
