python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench n8 rc=$?"; tail -c 300 gpurun_out/bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['result_equal'], d['roofline']['frac'], d['timing']['device_ms']['median'])
for k,v in d.get('workloads',{}).items(): print(k, v.get('ms_per_step'), v.get('result_equal'))
PY
