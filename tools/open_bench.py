"""Open path (SegmentManager upload): phase times and GB/s for the bench tables.  usage: tools/open_bench.py [rows] [kind]"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from immutable3_b200 import SegmentManager
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "pfor"
d, table, _ = bench.ensure_table(kind, rows, 0, 1, lambda: None, "product")
os.environ["IMM3_OPEN_TRACE"] = "1"
for threads in (None, 4, 8, 16, 32):
    if threads:
        os.environ["IMM3_IO_THREADS"] = str(threads)
    t0 = time.perf_counter()
    sm = SegmentManager(d)
    dt = time.perf_counter() - t0
    b = sm.getTable(table).resident_bytes
    print(f"threads={threads or 'default'} open_s={dt:.3f} resident={b/1e9:.2f} GB -> {b/dt/1e9:.1f} GB/s", flush=True)
    sm.close()
