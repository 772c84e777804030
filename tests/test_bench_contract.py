"""bench.py's reference arm (the CPU restatement of the reference path) prints the contract's JSON line; its `config`
is built by the same function as the GPU arm's, so the two lines describe one configuration.  CPU only: the GPU arm needs a B200."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, IMM3_BENCH_ALLOW_SHORT="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "2100000", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "rows/sec for scan+filter+project" and line["unit"] == "rows/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 1 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cfg = line["config"]
    assert cfg["rows_total"] == 2100000 and cfg["segments"] == 3 and cfg["scaling"] == "strong" and "C4" in cfg["workload"]
    assert cfg["result_rows"] == line["result_rows"] == 20999          # the 1 % id window, both ends exclusive


def test_both_arms_build_config_from_the_same_function():
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("protocol_keys(args,") == 3                       # the definition, reference_arm and the GPU arm
    assert src.count('"config": config_dict(args, args.workload, total, world, protocol_keys(') == 2
