#!/usr/bin/env python
"""Writer-side measurement: the sorted-integer encoder on the GPU (imm3_pfor_encode_blocks_gpu, whole call: H2D of the
values, size pass, encode pass, D2H of the bytes) next to the host encoder (imm3_pfor_encode, one block at a time, one
thread).  Usage: python tools/encode_bench.py [rows]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from immutable3_b200.loader import pfor_encode, pfor_encode_blocks_gpu  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
ids = np.arange(n, dtype=np.int32)
pfor_encode_blocks_gpu(ids[: 1 << 20])  # context creation, module load
t0 = time.perf_counter()
enc, off = pfor_encode_blocks_gpu(ids)
t_gpu = time.perf_counter() - t0
sample = min(n, 4_000_000)
t0 = time.perf_counter()
host = b"".join(pfor_encode(ids[i:i + 1024]) for i in range(0, sample, 1024))
t_cpu = (time.perf_counter() - t0) * n / sample
assert enc[: len(host)] == host
print(json.dumps({"rows": n, "encoded_bytes": len(enc), "gpu_call_s": t_gpu, "gpu_rows_per_s": n / t_gpu,
                  "host_1thread_s_extrapolated": t_cpu, "host_rows_per_s": n / t_cpu, "bytes_per_row": len(enc) / n}))
