python -m pytest tests -m gpu -x -q 2>&1 | tail -4
IMM3_BENCH_NO_STAGES=1 IMM3_DEBUG=16 IMM3_TRACE=gpurun_out/trace_ge3.txt python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-secondary --no-verify > gpurun_out/trace_ge3.log 2>&1
head -11 gpurun_out/trace_ge3.txt
tools/ab_c4.sh "IMM3_X=0" "IMM3_NO_GROUPEMIT=1"
AB_ARGS="--workload c4_limit10" tools/ab_c4.sh "IMM3_X=0"
