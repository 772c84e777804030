python -m pytest tests/test_gpu_agg.py -x -q 2>&1 | tail -5
for w in agg agg_count; do
python bench.py --workload $w --rows 100000000 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/agg_$w.json 2> gpurun_out/agg_$w.err
python - gpurun_out/agg_$w.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']; t=d['timing']
    print(d['config']['workload'][:60], "device", t['device_ms']['median'], "wall", t['wall_ms']['median'], "frac", r['frac'], "equal", d.get('result_equal'))
except Exception as e:
    print("FAILED", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
