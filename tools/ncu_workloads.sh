export IMM3_BENCH_ALLOW_SHORT=1
for w in ${WORKLOADS:-c3 c4 c5p_nolimit x_all x_1pct}; do
  CMD="python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
  $CMD > gpurun_out/plain_$w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"filter_kernel|emit_|blocks_|scan_dense|scan_blocks" -s 8 -c 3 -o gpurun_out/prof_r1_$w $CMD > gpurun_out/ncu_$w.log 2>&1
  tail -1 gpurun_out/ncu_$w.log
done
