python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 400 gpurun_out/bench_reference.json
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
IMM3_BENCH_NO_STAGES=1 tools/ncu_c4.sh r2final2 unpruned blocks_filter_lane
IMM3_BENCH_NO_STAGES=1 tools/ncu_c4.sh r2final2b unpruned blocks_group_emit
IMM3_BENCH_NO_STAGES=1 tools/ncu_c4.sh r2final2 pruned blocks_prune
ls gpurun_out/*r2final2*
