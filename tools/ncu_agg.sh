#!/bin/bash
# ncu captures of the aggregation query's kernels at 100 M rows (run under gpurun): launch list with durations + one full capture of agg_kernel.
# usage: tools/ncu_agg.sh <tag> [workload]
set -u
TAG=${1:-r2}
W=${2:-agg}
export IMM3_BENCH_ALLOW_SHORT=1
ARGS="--workload $W --rows 100000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-verify"
python bench.py $ARGS > gpurun_out/plain_${W}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${W}.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:'agg_|filter_kernel' --launch-skip 0 -c 30 --csv \
    --log-file gpurun_out/launches_${W}_${TAG}.csv python bench.py $ARGS > gpurun_out/ncu_l_${W}.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'agg_kernel' --launch-skip 3 -c 1 -o gpurun_out/prof_${W}_agg_kernel_${TAG} python bench.py $ARGS > gpurun_out/ncu_e_${W}.log 2>&1
echo "full capture rc=$?"
