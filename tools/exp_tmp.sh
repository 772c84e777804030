for v in "IMM3_FILTER_STAGES=4" "IMM3_FILTER_STAGES=2" "IMM3_FILTER_STAGES=6" "IMM3_FILTER_STAGES=8"; do
for w in x_count; do
env $v python bench.py --workload $w --rows 1000000000 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --no-verify > gpurun_out/ab_$w.json 2> gpurun_out/ab_$w.err
python - "$v $w" gpurun_out/ab_$w.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r=d['roofline']; t=d['timing']
    print(f"{sys.argv[1]:30s} device {t['device_ms']['median']*1e3:8.1f} us  wall {t['wall_ms']['median']*1e3:8.1f} us  frac {r['frac']:.3f} equal {d.get('result_equal')}")
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[2].replace('.json','.err')).read()[-800:])
PY
done
done
