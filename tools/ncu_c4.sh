#!/bin/bash
# ncu captures of the C4 query's kernels at 1 B rows (run under gpurun): launch list with durations + one full capture per kernel.
# usage: tools/ncu_c4.sh <tag> [pruned|unpruned] [kernel regex for the full capture]
set -u
TAG=${1:-r2}
MODE=${2:-unpruned}
KERN=${3:-blocks_filter}
if [ "$MODE" = "unpruned" ]; then export IMM3_NO_PRUNE=1; fi
export IMM3_BENCH_KEEP_PRUNE_ENV=1
ARGS="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-verify"
python bench.py $ARGS > gpurun_out/plain_c4_${MODE}.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:'blocks_|offset_scan|count_exchange' --launch-skip 0 -c 40 --csv \
    --log-file gpurun_out/launches_c4_${MODE}_${TAG}.csv python bench.py $ARGS > gpurun_out/ncu_l_${MODE}.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"$KERN" --launch-skip 3 -c 1 -o gpurun_out/prof_c4_${KERN}_${MODE}_${TAG} python bench.py $ARGS > gpurun_out/ncu_e_${MODE}.log 2>&1
echo "full capture rc=$?"
