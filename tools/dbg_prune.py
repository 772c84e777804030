import os, sys, time
sys.path.insert(0, os.getcwd())
import bench
rows = int(sys.argv[1]); flags = int(sys.argv[2]); use_torch = int(sys.argv[3])
if use_torch:
    import torch
    torch.cuda.set_device(0)
    buf = torch.zeros(512 << 20, dtype=torch.uint8, device="cuda")
from immutable3_b200 import SegmentManager, Engine, OPEN_KEEP_HOST
d, table, _ = bench.ensure_table("pfor", rows, 0, 1, lambda: None, "product")
sm = SegmentManager(d, flags=flags)
eng = Engine(sm)
q = bench.build_query("c4", table, rows)
prep = eng.prepare(q)
for i in range(4):
    try:
        if use_torch:
            buf.view(torch.int64).sum().item()
        r = eng.begin_prepared(prep)
        print("ok", i, rows, flags, r.local_count, r.kernel_launches)
        r.close()
    except Exception as e:
        print("FAIL", i, rows, flags, e)
sm.close()
