tools/ab_c4.sh "IMM3_DEBUG=0" "IMM3_DEBUG=768" "IMM3_DEBUG=0 IMM3_X=1" "IMM3_DEBUG=768 IMM3_X=1" "IMM3_DEBUG=256 IMM3_X=1"
IMM3_BENCH_NO_STAGES=1 IMM3_DEBUG=784 IMM3_TRACE=gpurun_out/trace_ge4.txt python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-secondary --no-verify > gpurun_out/trace_ge4.log 2>&1
head -11 gpurun_out/trace_ge4.txt
