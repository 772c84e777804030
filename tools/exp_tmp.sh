T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
$T bench.py --gpus 8 --sweep --steps 10 > gpurun_out/sweep_n8.json 2> gpurun_out/sweep_n8.err
tail -c 300 gpurun_out/sweep_n8.err
$T bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -c 300 gpurun_out/bench_n8.err
tail -c 1500 gpurun_out/bench_n8.json | head -c 600
