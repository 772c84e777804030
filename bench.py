#!/usr/bin/env python
"""bench.py — scan + filter + project throughput of the B200 path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c4dense] [--rows R]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # CPU restatement of the reference path on the host cores

A "step" is one pass of the hot path (one query) over the whole resident table.  Default workload is
BASELINE.json configs[1] (C2): test_100m, `select id, age where (age > 18 and age < 30)`, no limit.
Weak scaling: every rank owns a 100 M-row canonical slice of an (N x 100 M)-row table; the data path has
no collective, only the per-rank match counts are all-gathered (NCCL).

Printed keys (one JSON line on rank 0):
  value        rows/s, table resident in HBM, CUDA-event time of the query's kernels (filter -> offset scan -> emit) (max over ranks)
  e2e          rows/s through the C ABI from PINNED HOST buffers: per step the used columns are re-staged
               host->HBM, the query runs, counts are exchanged and the result rows are copied to host
  e2e_resident rows/s through the C ABI with the table resident (the steady state of the drop-in:
               SegmentManager loads once, Engine.execute per query): query + count exchange + result D2H
  roofline     achieved algorithmic GB/s of the query's kernels against the measured HBM peak
  cpu_baseline the oracle (CPU restatement, kind "port") timed on this box's host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK, SEGMENT = 1024, 1000
ROWS_PER_SEG = BLOCK * SEGMENT + 1

WORKLOADS = {
    # name: (description, id codec, predicates as (col, op, value), projection, limit)
    "c2": ("C2 test_100m: select id, age from test_100m where (age > 18 and age < 30)", "DENSE_INT"),
    "c3": ("C3 test_100m: select id, state, age from test_100m where (state = 'CA' and age > 18 and age < 30)", "DENSE_INT"),
    "c4": ("C4 sorted-int-codec id: select id from t where (id > L and id < H), 1% window", "PFOR_INT"),
    "c4dense": ("C4 dense-id twin: select id from t where (id > L and id < H), 1% window", "DENSE_INT"),
    # experiments (same test_100m table as c2): isolate the cost of each stage of the pipeline
    "x_age": ("X select age where (age > 18 and age < 30)", "DENSE_INT"),
    "x_id": ("X select id where (age > 18 and age < 30)", "DENSE_INT"),
    "x_count": ("X select <nothing> where (age > 18 and age < 30)", "DENSE_INT"),
    "x_none": ("X select id, age where age = 127 (no match)", "DENSE_INT"),
    "x_all": ("X select id, age (no predicate)", "DENSE_INT"),
    "x_1pct": ("X select id, age where age = 7", "DENSE_INT"),
    **{f"x_lt{t}": (f"C5 scattered: select id, age where age < {t} ({t} %)", "DENSE_INT") for t in (2, 3, 4, 5, 6, 8, 10, 25, 50)},
    "c5_limit10": ("C5 select id, age where (age > 18 and age < 30) limit 10", "DENSE_INT"),
    "c5_rare_limit10": ("C5 select id, state, age where (state = 'CA' and age = 7) limit 10", "DENSE_INT"),
    "c5p_limit10": ("C5 sorted-int-codec id: select id, age where age < 10 limit 10", "PFOR_INT"),
    "c5p_nolimit": ("C5 sorted-int-codec id: select id, age where age < 10", "PFOR_INT"),
    "c5p_window_limit10": ("C5 sorted-int-codec id: select id where (id > L and id < H) limit 10, 1% window", "PFOR_INT"),
}


def build_query(workload: str, table: str, total_rows: int):
    from immutable3_b200 import GT, LT, And, Match, Project, Query, Select

    if workload == "c2":
        return Query(table, And(Select("age", GT(18)), Select("age", LT(30))), Project(["id", "age"]))
    if workload.startswith("x_"):
        from immutable3_b200 import EQ, NoSelect
        sel = And(Select("age", GT(18)), Select("age", LT(30)))
        proj = {"x_age": ["age"], "x_id": ["id"], "x_count": []}.get(workload, ["id", "age"])
        if workload == "x_none":
            sel = Select("age", EQ(127))
        if workload == "x_all":
            sel = NoSelect
        if workload == "x_1pct":
            sel = Select("age", EQ(7))
        if workload.startswith("x_lt"):
            sel = Select("age", LT(int(workload[4:])))
        return Query(table, sel, Project(proj))
    if workload == "c5_limit10":
        return Query(table, And(Select("age", GT(18)), Select("age", LT(30))), Project(["id", "age"], 10))
    if workload in ("c5p_limit10", "c5p_nolimit"):
        return Query(table, Select("age", LT(10)), Project(["id", "age"], 10 if workload == "c5p_limit10" else 0))
    if workload == "c5p_window_limit10":
        lo, hi = total_rows // 2 - total_rows // 200, total_rows // 2 + total_rows // 200
        return Query(table, And(Select("id", GT(lo)), Select("id", LT(hi))), Project(["id"], 10))
    if workload == "c5_rare_limit10":
        from immutable3_b200 import EQ
        return Query(table, And(Select("state", Match(["CA"])), Select("age", EQ(7))), Project(["id", "state", "age"], 10))
    if workload == "c3":
        return Query(table, And(And(Select("state", Match(["CA"])), Select("age", GT(18))), Select("age", LT(30))),
                     Project(["id", "state", "age"]))
    lo, hi = total_rows // 2 - total_rows // 200, total_rows // 2 + total_rows // 200
    return Query(table, And(Select("id", GT(lo)), Select("id", LT(hi))), Project(["id"]))


def oracle_query_args(query):
    """The same query as oracle predicate tuples (col, op, value)."""
    from immutable3_b200 import EQ, GT, LT, Match, flatten_select

    out = []
    for leaf in flatten_select(query.select):
        c = leaf.cond
        if isinstance(c, GT):
            out.append((leaf.col, 1, c.gt))
        elif isinstance(c, LT):
            out.append((leaf.col, 2, c.lt))
        elif isinstance(c, EQ):
            out.append((leaf.col, 3, c.eq))
        elif isinstance(c, Match):
            out.append((leaf.col, 4, list(c.values)))
    return out, list(query.project.cols), int(query.project.limit)


def canonical_ids(nseg: int):
    return sorted(range(nseg), key=lambda i: f"id_{i}.dat")  # SegmentManager.scala:41


def data_dir_for(args, world: int) -> str:
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    kind = "c2" if args.workload in ("c2", "c3") or args.workload.startswith("x_") or args.workload.startswith("c5_") else args.workload
    return os.path.join(base, f"imm3_bench_{kind}_{args.rows}x{world}")


def ensure_table(args, rank: int, world: int, barrier):
    """Cooperatively write the synthetic table: rank 0 clears + writes _table.meta, then every rank
    writes the segment files of its own canonical slice.  Reused if already complete."""
    from immutable3_b200 import _lib as L
    from immutable3_b200.dist import shard_range
    from immutable3_b200.loader import synth_segments, synth_write

    total = args.rows * world
    d = data_dir_for(args, world)
    table = "test_100m" if args.workload in ("c2", "c3") or args.workload.startswith("x_") or args.workload.startswith("c5_") else "test_ids"
    marker = os.path.join(d, table, f".complete_{rank}_{world}")
    codec = L.CODEC_PFOR_INT if WORKLOADS[args.workload][1] == "PFOR_INT" else L.CODEC_DENSE_INT
    nseg = synth_segments(total, BLOCK, SEGMENT)
    have = os.path.exists(marker) or os.path.exists(os.path.join(d, table, ".complete_all"))
    t0 = time.time()
    if rank == 0 and not os.path.exists(os.path.join(d, table, "_table.meta")):
        os.makedirs(d, exist_ok=True)
        synth_write(d, table, total, BLOCK, SEGMENT, codec, 0, 0, True)
    barrier()
    if not have:
        a, b = shard_range(nseg, rank, world)
        for sid in canonical_ids(nseg)[a:b]:
            synth_write(d, table, total, BLOCK, SEGMENT, codec, sid, sid + 1, False)
        open(marker, "w").write("ok")
    barrier()
    return d, table, total, time.time() - t0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(workload: str):
    """dram__bytes_read+write per query (sum over its kernels) from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload)
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restatement of the reference's Scala path), one task per segment on a pool
# ---------------------------------------------------------------------------------------------------
def run_cpu(args, d, table, query, nthreads, steps, warmup, seg_begin=0, seg_end=-1):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import build_oracle

    build_oracle.build_oracle()
    import oracle_lib as O

    preds, proj, limit = oracle_query_args(query)
    times, nrows, rows_scanned = [], 0, 0
    with O.Oracle(d) as orc:
        rows_scanned = orc.nrows(table, seg_begin, seg_end)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            r = orc.query(table, preds, proj, limit=limit, nthreads=nthreads, seg_begin=seg_begin, seg_end=seg_end)
            dt = time.perf_counter() - t0
            nrows = r.nrows
            if i >= warmup:
                times.append(dt)
    return rows_scanned, nrows, times


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--rows", type=int, default=100_000_000, help="rows per GPU (weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-tma", action="store_true", help="direct-load variant of the dense kernel (A/B)")
    ap.add_argument("--cpu-steps", type=int, default=2)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours" and not os.environ.get("IMM3_BENCH_ALLOW_SHORT"):
        args.warmup = max(args.warmup, 3)  # timing rule: W >= 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    # ------------------------------------------------------------------ reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        from immutable3_b200 import _build

        _build.build_lib()  # the table writer lives in the product library (host-only code)
        nw = max(1, args.gpus)
        d, table, total, gen_s = ensure_table(args, 0, 1, lambda: None) if nw == 1 else ensure_table_single(args, nw)
        query = build_query(args.workload, table, total)
        cores = os.cpu_count() or 1
        steps = max(1, min(args.steps, 5))
        rows, nres, times = run_cpu(args, d, table, query, cores, steps, min(args.warmup, 1))
        sec = statistics.mean(times)
        v = rows / sec
        line = {
            "impl": "reference", "metric": "rows/sec for scan+filter+project", "value": v, "unit": "rows/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload][0], "rows": rows, "block_size": BLOCK, "segment_size": SEGMENT,
                       "note": "CPU restatement (oracle port) of the reference Scala path; the Scala engine itself cannot run (no JVM)"},
            "cpu_baseline": {"value": v, "unit": "rows/s", "cores": cores, "kind": "port",
                             "sample": f"whole table ({rows} rows), {steps} passes, one task per segment on {cores} threads"},
            "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "result_rows": nres,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ GPU arm
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the scan path has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        barrier = lambda: (torch.cuda.synchronize(), dist.barrier())  # noqa: E731
    else:
        barrier = lambda: torch.cuda.synchronize()  # noqa: E731

    from immutable3_b200 import OPEN_KEEP_HOST, OPEN_NO_TMA, Engine, SegmentManager, flatten_select
    from immutable3_b200.dist import limit_split

    d, table, total, gen_s = ensure_table(args, rank, world, barrier)
    query = build_query(args.workload, table, total)
    flags = OPEN_KEEP_HOST | (OPEN_NO_TMA if args.no_tma else 0)
    t0 = time.perf_counter()
    sm = SegmentManager(d, device=local_rank, rank=rank, world=world, flags=flags)
    open_s = time.perf_counter() - t0
    tinfo = sm.getTable(table)
    eng = Engine(sm)
    used_cols = sorted({l.col for l in flatten_select(query.select)} | set(query.project.cols))
    used_bytes = sum(c.encoded_bytes for c in tinfo.columns if c.name in used_cols)

    flush_buf = torch.zeros(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    flush_sink = torch.zeros((), dtype=torch.int64, device="cuda")

    flush_mode = os.environ.get("IMM3_BENCH_FLUSH", "read")

    def flush_l2(i):
        # Evict the table from the 126 MB L2 between steps.  A READ pass over 512 MiB leaves clean lines behind; a write
        # pass (fill_) would leave ~126 MB of dirty lines whose write-back is then charged to the next query.
        if flush_mode == "write":
            flush_buf.fill_(i & 0xFF)
        elif flush_mode == "read":
            flush_sink.copy_(flush_buf.view(torch.int64).sum())
        torch.cuda.synchronize()

    def exchange(count):
        if world == 1:
            return [count]
        mine = torch.tensor([count], dtype=torch.int64, device="cuda")
        allc = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)  # the only exchange on the path: 8 bytes per rank over NVLink
        return [int(t.item()) for t in allc]

    # ---- value: table resident, CUDA-event time of the query's kernels ----
    sampler = ClockSampler(local_rank)
    kernel_ms, alg_bytes, launches, local_rows = [], 0, 0, 0
    stage_ms = [[], []]
    for i in range(args.warmup):
        flush_l2(i)
        with eng.begin(query) as r:
            exchange(r.local_count)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush_l2(i)
        with eng.begin(query) as r:
            kernel_ms.append(r.device_ms)
            stage_ms[0].append(r.stage_ms(0))
            stage_ms[1].append(r.stage_ms(1))
            alg_bytes = r.algorithmic_bytes
            launches += r.kernel_launches
            local_rows = r.local_count
            counts = exchange(r.local_count)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    total_ms = sum(kernel_ms)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
        total_ms_max = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches_all = int(lt.item())
    else:
        total_ms_max, launches_all = total_ms, launches
    ms_per_step = total_ms_max / args.steps
    value = total / (ms_per_step * 1e-3)

    # ---- e2e through the C ABI ----
    def e2e_pass(reupload: bool):
        """K steps back to back, timed as a whole.  Every step stages its inputs (H2D), runs the query, exchanges the
        counts and reads its result rows back (D2H); the read-back of step i is asynchronous and overlaps the staging of
        step i+1 (full-duplex PCIe), and every result is waited for and checked before the clock stops."""
        w_e2e = max(2, args.warmup // 2)
        d2h = 0
        t0 = 0.0
        prev = None
        for i in range(w_e2e + args.steps):
            if i == w_e2e:
                if prev is not None:
                    prev.wait().close()
                    prev = None
                sm.sync()
                barrier()
                t0 = time.perf_counter()
            if reupload:
                sm.reupload(table, used_cols)          # H2D of this step's inputs from pinned host memory (async)
            r = eng.begin(query)                       # kernels, local count (returns when the count is known)
            cnts = exchange(r.local_count)
            _, takes = limit_split(cnts, int(query.project.limit))
            r.fetch_async(takes[rank])                 # result rows -> pinned host buffers, on the copy stream
            if prev is not None:
                prev.wait()
                d2h = sum(prev.col_width(c) for c in range(prev.ncols)) * prev.nrows
                prev.close()
            prev = r
        prev.wait()
        d2h = sum(prev.col_width(c) for c in range(prev.ncols)) * prev.nrows
        assert prev.nrows == takes[rank]
        prev.close()
        sm.sync()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        bb = torch.tensor([d2h], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(bb)
        return float(tt.item()) / args.steps, int(bb.item())

    e2e = e2e_res = None
    if not args.no_e2e:
        sec_res, d2h_b = e2e_pass(False)
        sec_cold, _ = e2e_pass(True)
        e2e = {"value": total / sec_cold, "unit": "rows/s", "h2d_bytes_per_step": used_bytes * world, "d2h_bytes_per_step": d2h_b,
               "ms_per_step": sec_cold * 1e3,
               "what": "per step: re-stage the query's columns host(pinned)->HBM, kernels, count exchange, result rows -> host (read-back of step i overlaps the staging of step i+1; K steps timed as a whole)"}
        e2e_res = {"value": total / sec_res, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h_b,
                   "ms_per_step": sec_res * 1e3,
                   "what": "table resident in HBM (SegmentManager loads once): kernels, count exchange, result rows -> host"}

    # ---- roofline of the dominant (only) kernel ----
    peak, peak_src = measured_peak()
    launch_ms = statistics.mean(kernel_ms)
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    st0, st1 = statistics.mean(stage_ms[0]), statistics.mean(stage_ms[1])
    multi = launches >= 2 * args.steps  # filter kernel + emit kernel(s) per query
    if WORKLOADS[args.workload][1] == "PFOR_INT":
        kname = ("blocks_filter_kernel -> blocks_emit_kernel (sorted-integer codec decoded warp-per-block; timed together)" if multi
                 else "scan_blocks_kernel")
    elif multi:
        kname = "filter_kernel -> emit_stream_kernel | emit_kernel (one query = one launch of each; timed together)"
    else:
        kname = "scan_dense_kernel"
    # achieved = algorithmic bytes of the query / CUDA-event time of ALL its kernels (conservative: the numerator is the
    # SURVEY.md 8d lower bound on traffic, the denominator includes every stage and the gaps between them)
    traffic = recorded_traffic(args.workload) if args.rows == 100_000_000 else None  # (the capture is of the 100 M-row table)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": kname,
                # the DRAM bytes the query really moves (ncu) over the same time: how close the kernels run to the HBM
                # roofline of their actual traffic (128-byte line fills make it ~3x the algorithmic bytes on C2)
                "traffic_gbs": (traffic / (launch_ms * 1e-3) / 1e9) if traffic else None,
                "traffic_frac": (traffic / (launch_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms_mean": launch_ms, "launch_ms_min": min(kernel_ms),
                "stage_ms_mean": [st0, st1],
                "stages": ({"note": "per-stage CUDA-event times need IMM3_NO_PDL=1 (the emit kernel is a programmatic dependent launch of the filter "
                                    "kernel, no event may sit between them); ncu shares are in profiles/",
                            "filter_kernel(+offset scan)": st0 if st1 > 0 else None, "emit kernels": st1 if st1 > 0 else None} if multi else None)}

    # ---- CPU baseline beside it (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rows, nres, times = run_cpu(args, d, table, query, cores, args.cpu_steps, 1)
        cpu = {"value": rows / statistics.mean(times), "unit": "rows/s", "cores": cores, "kind": "port",
               "sample": f"whole table ({rows} rows), {args.cpu_steps} passes after 1 warm-up, one task per segment on {cores} threads",
               "result_rows_equal": nres == local_rows}

    if rank == 0:
        line = {
            "metric": "rows/sec for scan+filter+project", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload][0], "rows_per_gpu": args.rows, "rows_total": total, "block_size": BLOCK,
                       "segment_size": SEGMENT, "segments": tinfo.nsegments, "l2": "flushed between steps (512 MiB %s pass)" % os.environ.get("IMM3_BENCH_FLUSH", "read"),
                       "kernel_variant": "direct" if args.no_tma else "tma", "result_rows_rank0": local_rows},
            "clocks": clocks, "e2e": e2e, "e2e_resident": e2e_res, "gpu_launches": launches_all, "roofline": roofline,
            "cpu_baseline": cpu, "wall_s_timed_region": wall, "open_s": open_s, "upload_gbs": tinfo.resident_bytes / open_s / 1e9,
            "table_gen_s": gen_s,
        }
        print(json.dumps(line))
    sm.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def ensure_table_single(args, nw):
    """Reference arm at N>1: one process writes the whole (N x rows) table."""
    class A:
        pass

    a = A()
    a.__dict__.update(vars(args))
    from immutable3_b200 import _lib as L
    from immutable3_b200.loader import synth_write

    total = args.rows * nw
    d = data_dir_for(args, nw)
    table = "test_100m" if args.workload in ("c2", "c3") else "test_ids"
    codec = L.CODEC_PFOR_INT if WORKLOADS[args.workload][1] == "PFOR_INT" else L.CODEC_DENSE_INT
    marker = os.path.join(d, table, ".complete_all")
    t0 = time.time()
    by_ranks = all(os.path.exists(os.path.join(d, table, f".complete_{r}_{nw}")) for r in range(nw))
    if not os.path.exists(marker) and not by_ranks:
        os.makedirs(d, exist_ok=True)
        synth_write(d, table, total, BLOCK, SEGMENT, codec, 0, -1, True)
        open(marker, "w").write("ok")
    return d, table, total, time.time() - t0


if __name__ == "__main__":
    sys.exit(main())
