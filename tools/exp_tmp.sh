python -m pytest tests/test_gpu_parity.py tests/test_gpu_real_or.py -x -q 2>&1 | tail -3
for w in c5_rare c5p_lt1 c5p_lt10 c2p; do
python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/ab_$w.json 2> gpurun_out/ab_$w.err
python - "$w" gpurun_out/ab_$w.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r=d['roofline']; t=d['timing']
    print(f"{sys.argv[1]:30s} device {t['device_ms']['median']*1e3:8.1f} us  wall {t['wall_ms']['median']*1e3:8.1f} us  frac {r['frac']:.3f} equal {d.get('result_equal')} stages {list((r.get('stage_ms_mean') or {}).values())[:2]}")
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[2].replace('.json','.err')).read()[-800:])
PY
done
