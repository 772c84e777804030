#!/bin/bash
# tools/gpu_sweep.sh — A/B sweep of bench.py workloads and kernel knobs on one B200 (run under gpurun).
# usage: tools/gpu_sweep.sh OUTFILE "<env assignments>|<bench args>" ...
out=$1; shift
mkdir -p "$(dirname "$out")"
: > "$out"
for spec in "$@"; do
  envs=${spec%%|*}; args=${spec#*|}
  echo "## $envs | $args" >> "$out"
  env $envs python bench.py --no-cpu-baseline --no-e2e --steps 20 --warmup 3 $args 2>&1 | tail -1 | python -c '
import json,sys
l=sys.stdin.read().strip()
try:
    j=json.loads(l); r=j["roofline"]
    print("ms=%.4f min=%.4f GB/s=%.0f frac=%.3f rows/s=%.3e res=%s stage=%s" % (j["ms_per_step"], r["launch_ms_min"], r["achieved"], r["frac"], j["value"], j["config"].get("result_rows_rank0"), r.get("stage_ms_mean")))
except Exception as e:
    print("FAILED", l[-400:])
' >> "$out"
done
cat "$out"
