"""Writer side: SegmentWriter + LoaderCli roll logic (Segment.scala:70-152, LoaderCli.scala:113-154)
and the deterministic synthetic tables of BASELINE.md — thin wrappers over the C ABI."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib as L


class SegmentWriter:
    """One writer per table; append typed column arrays or CSV lines, exactly the loader's layout
    (every full segment holds segment_size*block_size + 1 rows)."""

    def __init__(self, data_dir: str, table: str, col_specs: Sequence[str], block_size: int, segment_size: int,
                 first_segment_id: int = 0, write_table_meta: bool = True):
        self._lib = L.lib()
        self._h = C.c_void_p()
        self.col_specs = list(col_specs)
        specs = L.cstr_array(self.col_specs)
        L.check(self._lib.imm3_writer_open(str(data_dir).encode(), table.encode(), C.cast(specs, C.POINTER(C.c_char_p)),
                                           len(self.col_specs), block_size, segment_size, first_segment_id,
                                           1 if write_table_meta else 0, C.byref(self._h)))

    def append(self, *cols):
        """cols: one array per column — int32 for INT, int8 for TINYINT, 'S<k>' (or uint8 [n,k]) for STRING(k)."""
        assert len(cols) == len(self.col_specs)
        arrs = [np.ascontiguousarray(c) for c in cols]
        n = len(arrs[0])
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        L.check(self._lib.imm3_writer_append(self._h, ptrs, n))

    def append_csv_line(self, line: str):
        L.check(self._lib.imm3_writer_append_csv_line(self._h, line.encode()))

    def close(self):
        if self._h:
            h, self._h = self._h, None
            L.check(self._lib.imm3_writer_close(h))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def load_csv(data_dir: str, table: str, col_specs: Sequence[str], block_size: int, segment_size: int, csv_path: str):
    """LoaderCli main: `-t table -c specs -d data_dir -i csv --block-size B --segment-size S`."""
    lib = L.lib()
    specs = L.cstr_array(list(col_specs))
    L.check(lib.imm3_load_csv(str(data_dir).encode(), table.encode(), C.cast(specs, C.POINTER(C.c_char_p)), len(col_specs),
                              block_size, segment_size, str(csv_path).encode()))


def pfor_encode(values) -> bytes:
    """PFORCodecInt.encode of one block (PFORCodec.scala:17-28)."""
    lib = L.lib()
    v = np.ascontiguousarray(values, dtype=np.int32)
    p = v.ctypes.data_as(C.POINTER(C.c_int32))
    n = L.check(lib.imm3_pfor_encode(p, len(v), None, 0))
    out = (C.c_uint8 * n)()
    L.check(lib.imm3_pfor_encode(p, len(v), out, n))
    return bytes(out)


def pfor_encode_blocks_gpu(values, block_rows: int = 1024, device: int = 0):
    """PFORCodecInt.encode of a whole column on the GPU: (bytes of all blocks back to back, int64 block offsets)."""
    lib = L.lib()
    v = np.ascontiguousarray(values, dtype=np.int32)
    p = v.ctypes.data_as(C.POINTER(C.c_int32))
    nblocks = (len(v) + block_rows - 1) // block_rows
    off = np.zeros(nblocks + 1, dtype=np.int64)
    po = off.ctypes.data_as(C.POINTER(C.c_int64))
    n = L.check(lib.imm3_pfor_encode_blocks_gpu(device, p, len(v), block_rows, None, 0, po))
    out = np.zeros(max(n, 1), dtype=np.uint8)
    L.check(lib.imm3_pfor_encode_blocks_gpu(device, p, len(v), block_rows, out.ctypes.data_as(C.POINTER(C.c_uint8)), n, po))
    return out[:n].tobytes(), off


def synth_segments(nrows: int, block_size: int, segment_size: int) -> int:
    rows_per_seg = block_size * segment_size + 1
    return (nrows + rows_per_seg - 1) // rows_per_seg


def synth_write(data_dir: str, table: str, nrows: int, block_size: int = 1024, segment_size: int = 1000,
                id_codec: int = L.CODEC_DENSE_INT, seg_id_begin: int = 0, seg_id_end: int = -1,
                write_table_meta: bool = True):
    """Write segments [seg_id_begin, seg_id_end) (numeric ids) of the synthetic table.  With
    write_table_meta the table directory is cleared first — do that on one rank, before the others write."""
    L.check(L.lib().imm3_synth_write(str(data_dir).encode(), table.encode(), nrows, block_size, segment_size, id_codec,
                                     seg_id_begin, seg_id_end, 1 if write_table_meta else 0))


def synth_rows(start: int, n: int):
    """(id, age, state) arrays of rows [start, start+n) of the synthetic generator."""
    lib = L.lib()
    ids = np.empty(n, np.int32)
    ages = np.empty(n, np.int8)
    states = np.empty(n, "S2")
    i, a, s = C.c_int32(), C.c_int8(), C.create_string_buffer(2)
    for k in range(n):
        lib.imm3_synth_row(start + k, C.byref(i), C.byref(a), s)
        ids[k], ages[k], states[k] = i.value, a.value, s.raw
    return ids, ages, states
