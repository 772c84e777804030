"""Shared helpers for the test-suite: building tables through the product writer, converting
between the Python Query ADT and the oracle's predicate tuples."""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np

import oracle_lib as O
from immutable3_b200 import And, EQ, GT, LT, Match, NoSelect, Or, Project, Query, Select, flatten_select
from immutable3_b200 import _lib as L
from immutable3_b200.loader import SegmentWriter

STATES = ["AL", "AK", "AZ", "AR", "CA", "CO", "CT", "DE", "FL", "GA", "HI", "ID", "IL", "IN", "IA", "KS", "KY",
          "LA", "ME", "MD", "MA", "MI", "MN", "MS", "MO", "MT", "NE", "NV", "NH", "NJ", "NM", "NY", "NC", "ND",
          "OH", "OK", "OR", "PA", "RI", "SC", "SD", "TN", "TX", "UT", "VT", "VA", "WA", "WV", "WI", "WY", "DC"]


def oracle_preds(select) -> list:
    out = []
    for leaf in flatten_select(select):
        c = leaf.cond
        if isinstance(c, GT):
            out.append((leaf.col, O.OP_GT, c.gt))
        elif isinstance(c, LT):
            out.append((leaf.col, O.OP_LT, c.lt))
        elif isinstance(c, EQ):
            out.append((leaf.col, O.OP_EQ, c.eq))
        elif isinstance(c, Match):
            out.append((leaf.col, O.OP_MATCH, list(c.values)))
        else:
            out.append((leaf.col, O.OP_NOOP, None))
    return out


def conj(*leaves):
    if not leaves:
        return NoSelect
    acc = leaves[0]
    for l in leaves[1:]:
        acc = And(acc, l)
    return acc


def make_table(data_dir, name, nrows, block_size, segment_size, seed=0, id_codec="DENSE_INT", extra_cols=(), id_mode="sorted"):
    """id:INT, state:STRING(2), age:TINYINT (+ optional extra columns) written with the product writer
    in the loader's layout.  Returns the column arrays."""
    rng = np.random.default_rng(seed)
    if id_mode == "sorted":
        ids = np.arange(nrows, dtype=np.int64) * 3 + 5
    elif id_mode == "steps":  # sorted with occasional big jumps and repeats
        ids = np.cumsum(rng.choice([0, 1, 2, 1000, 100000], size=nrows, p=[0.2, 0.5, 0.2, 0.09, 0.01]))
    else:  # unsorted, full int32 range incl. extremes
        ids = rng.integers(-2**31, 2**31, size=nrows, dtype=np.int64)
        if nrows > 4:
            ids[0], ids[1], ids[2], ids[3] = -2**31, 2**31 - 1, -1, 0
    ids = ids.astype(np.int64).astype(np.int32)
    ages = rng.integers(-128, 128, size=nrows, dtype=np.int64).astype(np.int8) if seed % 2 else rng.integers(0, 100, size=nrows, dtype=np.int64).astype(np.int8)
    states = np.array(STATES, dtype="S2")[rng.integers(0, len(STATES), size=nrows)]
    cols = {"id": ids, "state": states, "age": ages}
    specs = [f"id:{id_codec}", "state:DENSE_STRING:size=2", "age:DENSE_TINYINT"]
    for spec, arr in extra_cols:
        specs.append(spec)
        cols[spec.split(":")[0]] = arr
    with SegmentWriter(data_dir, name, specs, block_size, segment_size) as w:
        step = 100003
        names = [s.split(":")[0] for s in specs]
        for a in range(0, nrows, step):
            w.append(*[cols[n][a:a + step] for n in names])
        if nrows == 0:
            pass
    return cols


def numpy_expected(cols, select, proj, limit=0):
    """Independent numpy evaluation of the intended semantics over flat canonical-order arrays
    (only valid when the table's canonical order equals write order, i.e. < 11 segments)."""
    n = len(next(iter(cols.values())))
    m = np.ones(n, bool)
    for leaf in flatten_select(select):
        c, col = leaf.cond, cols[leaf.col]
        if isinstance(c, Match):
            m &= np.isin(col, np.array([v.encode() for v in c.values if len(v.encode()) == col.dtype.itemsize], dtype=col.dtype)) if any(
                len(v.encode()) == col.dtype.itemsize for v in c.values) else np.zeros(n, bool)
        else:
            v = c.gt if isinstance(c, GT) else c.lt if isinstance(c, LT) else c.eq
            k = O.lib().orc_d2i(float(v)) if col.dtype == np.int32 else O.lib().orc_d2b(float(v))
            m &= (col > k) if isinstance(c, GT) else (col < k) if isinstance(c, LT) else (col == k)
    idx = np.flatnonzero(m)
    if limit > 0:
        idx = idx[:limit]
    return [cols[p][idx] for p in proj], int(m.sum())
