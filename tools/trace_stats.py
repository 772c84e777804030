#!/usr/bin/env python
"""Summarise an IMM3_TRACE dump: per-tile globaltimer stamps [ticket, full, agg, prefix_seen, done, scanner_wrote]."""
import sys
import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8).astype(np.int64)
t0 = a[:, 0][a[:, 0] > 0].min()
names = ["ticket", "full", "agg", "prefix_seen", "done", "scan_wrote"]
rel = np.where(a[:, :6] > 0, a[:, :6] - t0, -1) / 1000.0
n = len(a)
print("tiles", n, "span_us", rel.max())
for i, nm in enumerate(names):
    v = rel[:, i]
    v = v[v >= 0]
    if len(v):
        print(f"{nm:12s} first={v.min():8.2f} median={np.median(v):8.2f} last={v.max():8.2f}")
def d(i, j, label):
    ok = (a[:, i] > 0) & (a[:, j] > 0)
    x = (a[ok, j] - a[ok, i]) / 1000.0
    if len(x):
        print(f"{label:28s} mean={x.mean():7.2f} p50={np.median(x):7.2f} p90={np.percentile(x,90):7.2f} max={x.max():7.2f}")
d(0, 1, "ticket -> full (TMA)")
d(1, 2, "full -> agg (filter)")
d(2, 5, "agg -> scanner wrote")
d(5, 3, "scanner wrote -> seen")
d(2, 3, "agg -> prefix seen")
d(3, 4, "prefix seen -> done (emit)")
d(1, 4, "full -> done (iteration)")
# scanner progress: time of scan_wrote by tile index in steps
sw = rel[:, 5]
idx = np.arange(n)
for q in (0, n // 8, n // 4, n // 2, 3 * n // 4, n - 1):
    print("tile", q, " ".join(f"{nm}={rel[q, i]:.2f}" for i, nm in enumerate(names)))
