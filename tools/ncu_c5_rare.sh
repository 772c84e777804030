export IMM3_BENCH_ALLOW_SHORT=1
ARGS="--workload c5_rare --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --no-verify"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --launch-skip 0 -c 40 --csv \
    --log-file gpurun_out/launches_c5_rare_r2y.csv python bench.py $ARGS > gpurun_out/ncu_l_c5_rare.log 2>&1
echo "launch list rc=$?"
