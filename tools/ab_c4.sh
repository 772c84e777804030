#!/bin/bash
# A/B of the C4 headline under environment switches (run under gpurun): tools/ab_c4.sh "VAR=1 VAR2=x" "..." ...
# prints device / wall / stage times per variant
for v in "$@"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary $AB_ARGS > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$v" gpurun_out/ab_$tag.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r=d['roofline']; t=d['timing']
    print(f"{sys.argv[1]:40s} device {t['device_ms']['median']*1e3:7.1f} us  wall {t['wall_ms']['median']*1e3:7.1f} us  frac {r['frac']:.3f}  stages {r.get('stage_ms_mean')}  equal {d.get('result_equal')}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
