python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "bench n4 rc=$?"; tail -c 300 gpurun_out/bench_n4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['result_equal'], d['roofline']['frac'], d['timing']['device_ms']['median'])
for k,v in d.get('workloads',{}).items(): print(k, v.get('ms_per_step'), v.get('result_equal'))
PY
