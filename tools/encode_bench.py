#!/usr/bin/env python
"""Writer-side measurement of the sorted-integer encoder on the GPU (imm3_pfor_encode_blocks_gpu):
  * device-resident values and output (how a table generated or loaded on the device is encoded): size pass + host scan of
    the block sizes + encode pass, wall clock around the synchronous call;
  * host values and output (H2D of the values and D2H of the bytes inside the call, pageable numpy memory);
  * the host encoder (imm3_pfor_encode, one block at a time, one thread) on a sample, extrapolated.
Usage: python tools/encode_bench.py [rows]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from immutable3_b200 import _lib as L  # noqa: E402
from immutable3_b200.loader import pfor_encode, pfor_encode_blocks_gpu  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
lib = L.lib()
ids = np.arange(n, dtype=np.int32)
pfor_encode_blocks_gpu(ids[: 1 << 20])  # context creation, module load

d_ids = torch.arange(n, dtype=torch.int32, device="cuda")
nblocks = (n + 1023) // 1024
off = np.zeros(nblocks + 1, dtype=np.int64)
po = off.ctypes.data_as(C.POINTER(C.c_int64))
pv = C.cast(d_ids.data_ptr(), C.POINTER(C.c_int32))
need = L.check(lib.imm3_pfor_encode_blocks_gpu(0, pv, n, 1024, None, 0, po))
d_out = torch.empty(need, dtype=torch.uint8, device="cuda")
pout = C.cast(d_out.data_ptr(), C.POINTER(C.c_uint8))
times = []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    L.check(lib.imm3_pfor_encode_blocks_gpu(0, pv, n, 1024, pout, need, po))
    times.append(time.perf_counter() - t0)
t_dev = min(times[1:])

t0 = time.perf_counter()
enc, off2 = pfor_encode_blocks_gpu(ids)
t_host_io = time.perf_counter() - t0
assert bytes(d_out.cpu().numpy().tobytes()) == enc and np.array_equal(off, off2)

sample = min(n, 4_000_000)
t0 = time.perf_counter()
host = b"".join(pfor_encode(ids[i:i + 1024]) for i in range(0, sample, 1024))
t_cpu = (time.perf_counter() - t0) * n / sample
assert enc[: len(host)] == host
print(json.dumps({"rows": n, "encoded_bytes": len(enc), "bytes_per_row": len(enc) / n,
                  "gpu_device_resident_s": t_dev, "gpu_device_resident_rows_per_s": n / t_dev,
                  "gpu_device_resident_GBps_in": 4 * n / t_dev / 1e9,
                  "gpu_host_io_s (sizing call + encode call, pageable H2D twice)": t_host_io,
                  "host_1thread_s_extrapolated": t_cpu, "host_1thread_rows_per_s": n / t_cpu}))
