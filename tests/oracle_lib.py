"""ctypes wrapper of oracle/liborc.so — the CPU restatement of the reference path.
TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "liborc.so")

OP_GT, OP_LT, OP_EQ, OP_MATCH, OP_NOTMATCH, OP_NOOP = 1, 2, 3, 4, 5, 6
COL_INT, COL_TINYINT, COL_STRING = 0, 1, 2


class OrcAgg(C.Structure):
    _fields_ = [("col", C.c_char_p), ("op", C.c_int32)]


AGG_COUNT, AGG_MIN, AGG_MAX = 0, 1, 2
COL_COUNT, COL_DOUBLE = 3, 4


class OrcPred(C.Structure):
    _fields_ = [("col", C.c_char_p), ("op", C.c_int32), ("num", C.c_double), ("strs", C.POINTER(C.c_char_p)), ("nstrs", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(LIB_PATH)
        P = C.c_void_p
        sig = {
            "orc_bytes_to_int": (C.c_int32, [C.c_char_p]),
            "orc_int_to_bytes": (None, [C.c_int32, C.c_char_p]),
            "orc_d2i": (C.c_int32, [C.c_double]),
            "orc_d2b": (C.c_int8, [C.c_double]),
            "orc_dense_decode": (C.c_int64, [C.c_char_p, C.c_int64, C.c_int, C.c_char_p]),
            "orc_iic_compress": (C.c_int64, [P, C.c_int32, P, C.c_int64]),
            "orc_iic_uncompress": (C.c_int32, [P, C.c_int64, P, C.c_int32]),
            "orc_pfor_encode_block": (C.c_int64, [P, C.c_int32, P, C.c_int64]),
            "orc_pfor_decode_block": (C.c_int32, [C.c_char_p, C.c_int64, P, C.c_int32]),
            "orc_open": (C.c_int, [C.c_char_p, C.POINTER(P)]),
            "orc_close": (None, [P]),
            "orc_table_nsegments": (C.c_int, [P, C.c_char_p]),
            "orc_table_block_size": (C.c_int, [P, C.c_char_p]),
            "orc_table_ncols": (C.c_int, [P, C.c_char_p]),
            "orc_segment_file_id": (C.c_int, [P, C.c_char_p, C.c_int]),
            "orc_table_nrows": (C.c_int64, [P, C.c_char_p, C.c_int, C.c_int]),
            "orc_query": (C.c_int, [P, C.c_char_p, C.POINTER(OrcPred), C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(P)]),
            "orc_query_agg": (C.c_int, [P, C.c_char_p, C.POINTER(OrcPred), C.c_int, C.POINTER(OrcAgg), C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.POINTER(P)]),
            "orc_result_nrows": (C.c_int64, [P]),
            "orc_result_nmatched": (C.c_int64, [P]),
            "orc_result_ncols": (C.c_int, [P]),
            "orc_result_col_type": (C.c_int, [P, C.c_int]),
            "orc_result_col_width": (C.c_int, [P, C.c_int]),
            "orc_result_col_data": (P, [P, C.c_int]),
            "orc_result_ref_throw": (C.c_int, [P, C.POINTER(C.c_int64)]),
            "orc_result_format_row": (C.c_int, [P, C.c_int64, C.c_char_p, C.c_size_t]),
            "orc_result_free": (None, [P]),
            "orc_filter_bitmap": (C.c_int, [P, C.c_char_p, C.POINTER(OrcPred), C.c_int, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
            "orc_free": (None, [P]),
            "orc_write_table_meta": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p]),
            "orc_write_column": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, P, C.c_int64, C.c_int, C.c_int]),
            "orc_synth_row": (None, [C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int8), C.c_char_p]),
            "orc_synth_write": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int, C.c_int]),
            "orc_last_error": (C.c_char_p, []),
        }
        for n, (r, a) in sig.items():
            f = getattr(l, n)
            f.restype, f.argtypes = r, a
        _lib = l
    return _lib


class OracleError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"[oracle status {status}] {msg}")
        self.status = status


def _check(rc):
    if rc < 0:
        raise OracleError(rc, (lib().orc_last_error() or b"").decode())
    return rc


def _preds(preds):
    """preds: list of (col, op, value) with value a number or a list of strings."""
    n = len(preds)
    arr = (OrcPred * max(1, n))()
    keep = []
    for i, (col, op, val) in enumerate(preds):
        arr[i].col = col.encode()
        arr[i].op = op
        if op in (OP_MATCH, OP_NOTMATCH):
            strs = (C.c_char_p * max(1, len(val)))(*[s.encode() for s in val])
            keep.append(strs)
            arr[i].strs = C.cast(strs, C.POINTER(C.c_char_p))
            arr[i].nstrs = len(val)
        elif val is not None:
            arr[i].num = float(val)
    return arr, n, keep


_NP = {COL_INT: np.dtype("<i4"), COL_TINYINT: np.dtype("i1"), COL_COUNT: np.dtype("<i8"), COL_DOUBLE: np.dtype("<f8")}


class OracleResult:
    def __init__(self, columns, types, widths, nmatched, ref_throw, ref_rows, rows_text):
        self.columns, self.types, self.widths = columns, types, widths
        self.nrows = len(columns[0]) if columns else 0
        self.nmatched, self.ref_throw, self.ref_rows = nmatched, ref_throw, ref_rows
        self._rows_text = rows_text

    def format_rows(self):
        return self._rows_text


class Oracle:
    def __init__(self, data_dir: str):
        self._l = lib()
        self._h = C.c_void_p()
        _check(self._l.orc_open(str(data_dir).encode(), C.byref(self._h)))

    def close(self):
        if self._h:
            self._l.orc_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def nsegments(self, table):
        return _check(self._l.orc_table_nsegments(self._h, table.encode()))

    def segment_file_ids(self, table):
        return [self._l.orc_segment_file_id(self._h, table.encode(), i) for i in range(self.nsegments(table))]

    def nrows(self, table, seg_begin=0, seg_end=-1):
        return _check(self._l.orc_table_nrows(self._h, table.encode(), seg_begin, seg_end))

    def query(self, table, preds, proj, limit=0, nthreads=1, seg_begin=0, seg_end=-1, fmt_rows=0) -> OracleResult:
        arr, n, keep = _preds(preds)
        pj = (C.c_char_p * max(1, len(proj)))(*[p.encode() for p in proj])
        out = C.c_void_p()
        _check(self._l.orc_query(self._h, table.encode(), arr, n, C.cast(pj, C.POINTER(C.c_char_p)), len(proj), limit, nthreads,
                                 seg_begin, seg_end, C.byref(out)))
        try:
            nrows = self._l.orc_result_nrows(out)
            cols, types, widths = [], [], []
            for c in range(self._l.orc_result_ncols(out)):
                t, w = self._l.orc_result_col_type(out, c), self._l.orc_result_col_width(out, c)
                dt = _NP.get(t, np.dtype(f"S{w}"))
                if nrows:
                    raw = (C.c_uint8 * (nrows * w)).from_address(self._l.orc_result_col_data(out, c))
                    cols.append(np.frombuffer(raw, dtype=dt, count=nrows).copy())
                else:
                    cols.append(np.empty(0, dt))
                types.append(t)
                widths.append(w)
            rb = C.c_int64()
            thr = self._l.orc_result_ref_throw(out, C.byref(rb))
            text = []
            buf = C.create_string_buffer(4096)
            for i in range(min(fmt_rows, nrows)):
                self._l.orc_result_format_row(out, i, buf, len(buf))
                text.append(buf.value.decode())
            res = OracleResult(cols, types, widths, self._l.orc_result_nmatched(out), thr, rb.value, text)
            if not cols:
                res.nrows = nrows
            return res
        finally:
            self._l.orc_result_free(out)

    def query_agg(self, table, preds, aggs, group_by=(), nthreads=1, seg_begin=0, seg_end=-1) -> OracleResult:
        """aggs: list of (op, col) with op in AGG_COUNT / AGG_MIN / AGG_MAX.  Columns of the result: the group columns, then
        one per aggregate (count: int64, min / max: float64), one row per group in first-appearance order."""
        arr, n, keep = _preds(preds)
        ag = (OrcAgg * max(1, len(aggs)))()
        for i, (op, col) in enumerate(aggs):
            ag[i].op, ag[i].col = op, col.encode()
        gb = (C.c_char_p * max(1, len(group_by)))(*[g.encode() for g in group_by])
        out = C.c_void_p()
        _check(self._l.orc_query_agg(self._h, table.encode(), arr, n, ag, len(aggs), C.cast(gb, C.POINTER(C.c_char_p)), len(group_by), nthreads,
                                     seg_begin, seg_end, C.byref(out)))
        try:
            nrows = self._l.orc_result_nrows(out)
            cols, types, widths = [], [], []
            for c in range(self._l.orc_result_ncols(out)):
                t, w = self._l.orc_result_col_type(out, c), self._l.orc_result_col_width(out, c)
                dt = _NP.get(t, np.dtype(f"S{w}"))
                if nrows:
                    raw = (C.c_uint8 * (nrows * w)).from_address(self._l.orc_result_col_data(out, c))
                    cols.append(np.frombuffer(raw, dtype=dt, count=nrows).copy())
                else:
                    cols.append(np.empty(0, dt))
                types.append(t)
                widths.append(w)
            res = OracleResult(cols, types, widths, self._l.orc_result_nmatched(out), 0, 0, [])
            res.nrows = nrows
            return res
        finally:
            self._l.orc_result_free(out)

    def filter_bitmap(self, table, preds, seg_begin=0, seg_end=-1):
        arr, n, keep = _preds(preds)
        words = C.POINTER(C.c_uint32)()
        nw, ns = C.c_int64(), C.c_int64()
        _check(self._l.orc_filter_bitmap(self._h, table.encode(), arr, n, seg_begin, seg_end, C.byref(words), C.byref(nw), C.byref(ns)))
        a = np.ctypeslib.as_array(words, shape=(max(1, nw.value),))[: nw.value].copy()
        self._l.orc_free(words)
        return a, ns.value


# value-level helpers -------------------------------------------------------------------------------
def pfor_encode(values) -> bytes:
    v = np.ascontiguousarray(values, dtype=np.int32)
    cap = 4 * (len(v) + len(v) // 32 + 64) + 8
    out = (C.c_uint8 * cap)()
    n = lib().orc_pfor_encode_block(v.ctypes.data, len(v), out, cap)
    assert n > 0
    return bytes(out[:n])


def pfor_decode(data: bytes, cap: int = 1 << 20) -> np.ndarray:
    out = np.empty(cap, np.int32)
    n = lib().orc_pfor_decode_block(data, len(data), out.ctypes.data, cap)
    if n < 0:
        raise OracleError(n, "bad PFOR block")
    return out[:n].copy()


def iic_compress(values) -> np.ndarray:
    v = np.ascontiguousarray(values, dtype=np.int32)
    cap = len(v) + len(v) // 32 + 64
    out = np.empty(cap, np.int32)
    n = lib().orc_iic_compress(v.ctypes.data, len(v), out.ctypes.data, cap)
    assert n > 0
    return out[:n].copy()


def dense_decode(data: bytes, width: int) -> bytes:
    out = C.create_string_buffer((len(data) // width + 2) * width)
    n = lib().orc_dense_decode(data, len(data), width, out)
    return out.raw[: n * width]


# writer side (fixtures / reference-arm tables without product code) ---------------------------------
CODEC_PFOR_INT, CODEC_DENSE_INT, CODEC_DENSE_TINYINT, CODEC_DENSE_STRING = 0, 1, 2, 3
_CODECS = {"PFOR_INT": (CODEC_PFOR_INT, "INT", 4), "DENSE_INT": (CODEC_DENSE_INT, "INT", 4), "DENSE_TINYINT": (CODEC_DENSE_TINYINT, "TINYINT", 1),
           "DENSE_STRING": (CODEC_DENSE_STRING, "STRING", 0)}


def write_table(data_dir, name, specs, cols, block_size, segment_size):
    """Oracle-side SegmentWriter + LoaderCli layout: specs use the loader's "name:CODEC[:k=v;k=v]" syntax
    (LoaderCli.scala:66-80), cols = {name: numpy array (int32 / int8 / S<k>)}."""
    import json

    metas, jobs = [], []
    for spec in specs:
        parts = spec.split(":")
        cname, codec = parts[0], parts[1]
        attrs = dict(kv.split("=") for kv in parts[2].split(";")) if len(parts) > 2 else {}
        cid, ctype, width = _CODECS[codec]
        if codec == "DENSE_STRING":
            width = int(attrs["size"])
        metas.append({"name": cname, "columnType": ctype, "codec": codec, "dtypeAttrs": attrs})
        arr = cols[cname]
        dt = {4: np.dtype("<i4"), 1: np.dtype("i1")}.get(width) if ctype != "STRING" else np.dtype(f"S{width}")
        jobs.append((cname, cid, width, np.ascontiguousarray(arr, dtype=dt)))
    meta = json.dumps({"name": name, "columns": metas, "blockSize": block_size}, separators=(",", ":"))
    _check(lib().orc_write_table_meta(str(data_dir).encode(), name.encode(), meta.encode()))
    for cname, cid, width, arr in jobs:
        _check(lib().orc_write_column(str(data_dir).encode(), name.encode(), cname.encode(), cid, width, arr.ctypes.data, len(arr), block_size, segment_size))


def synth_write(data_dir, table, nrows, block_size=1024, segment_size=1000, id_codec=CODEC_DENSE_INT, seg_begin=0, seg_end=-1,
                write_table_meta=True, nthreads=1):
    _check(lib().orc_synth_write(str(data_dir).encode(), table.encode(), nrows, block_size, segment_size, id_codec, seg_begin, seg_end,
                                 1 if write_table_meta else 0, nthreads))
