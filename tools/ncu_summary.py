#!/usr/bin/env python
"""Summarise ncu captures into small tracked files under profiles/ (runs on the CPU: `ncu -i` only reads the reports).
usage: python tools/ncu_summary.py <tag> <report.ncu-rep> [...]     -> profiles/<tag>_kernels_summary.csv + <tag>_<kernel>_raw.csv
       python tools/ncu_summary.py --launches <tag> <launches.csv>  -> profiles/<tag>_launches.csv (per launch: time, DRAM bytes, instructions)"""
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_active.avg", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    if sys.argv[1] == "--launches":
        tag, src = sys.argv[2], sys.argv[3]
        rows = [r for r in csv.reader(open(src)) if len(r) > 5]
        h = rows[0]
        ki, mi, vi, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
        d = {}
        for r in rows[1:]:
            d.setdefault((int(r[idi]), re.sub(r"\(.*", "", r[ki])), {})[r[mi]] = r[vi]
        dst = os.path.join(ROOT, "profiles", f"{tag}_launches.csv")
        with open(dst, "w") as f:
            f.write("id,kernel,gpu__time_duration.sum [ns],dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum\n")
            for (i, k), m in sorted(d.items()):
                f.write(f"{i},{k},{m.get('gpu__time_duration.sum')},{m.get('dram__bytes_read.sum')},{m.get('dram__bytes_write.sum')},{m.get('smsp__inst_executed.sum')}\n")
        print(dst)
        return
    tag = sys.argv[1]
    summ = os.path.join(ROOT, "profiles", f"{tag}_kernels_summary.csv")
    with open(summ, "w") as f:
        w = csv.writer(f)
        w.writerow(["report", "kernel"] + KEYS + ["stall reasons (warps per issue-active, > 0.15)"])
        for rep in sys.argv[2:]:
            h, units, data = raw_page(rep)
            for row in data:
                d = dict(zip(h, row))
                u = dict(zip(h, units))
                kname = re.sub(r"\(.*", "", d.get("Kernel Name", ""))
                stalls = []
                for k in h:
                    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                        try:
                            if float(d[k]) > 0.15:
                                stalls.append(k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "") + "=" + f"{float(d[k]):.2f}")
                        except ValueError:
                            pass
                w.writerow([os.path.basename(rep), kname] + [f"{d.get(k, '')} {u.get(k, '')}".strip() for k in KEYS] + [" ".join(stalls)])
                short = re.sub(r"[^A-Za-z0-9_]", "_", kname.replace("void ", ""))[:40]
                with open(os.path.join(ROOT, "profiles", f"{tag}_{short}_raw.csv"), "w") as g:
                    gw = csv.writer(g)
                    gw.writerow(["metric", "unit", "value"])
                    for k in h:
                        gw.writerow([k, u.get(k, ""), d.get(k, "")])
    print(summ)


if __name__ == "__main__":
    main()
