"""The oracle's restatement of the ProjectAgg branch (ProjectAggregate.scala:115-226, ProjectAggregateQueue.scala:9-54) against an
independent numpy evaluation: one row per group in first-appearance canonical order, count / min / max per group."""
import numpy as np
import pytest

import oracle_lib as O
from helpers import make_table


def _numpy_agg(cols, mask, aggs, group_by):
    idx = np.flatnonzero(mask)
    keys = list(zip(*[cols[g][idx] for g in group_by])) if group_by else [()] * len(idx)
    order, groups = [], {}
    for k, r in zip(keys, idx):
        if k not in groups:
            groups[k] = []
            order.append(k)
        groups[k].append(r)
    out = []
    for k in order:
        rows = np.array(groups[k])
        vals = []
        for op, col in aggs:
            if op == O.AGG_COUNT:
                vals.append(len(rows))
            elif op == O.AGG_MIN:
                vals.append(float(cols[col][rows].min()))
            else:
                vals.append(float(cols[col][rows].max()))
        out.append((k, vals))
    return out


@pytest.mark.parametrize("nrows,B,S", [(3000, 64, 5), (321, 64, 5), (1, 8, 2)])
def test_oracle_aggregation_matches_numpy(tmp_path, nrows, B, S):
    cols = make_table(tmp_path, "t", nrows, B, S, seed=3, id_mode="random")   # < 11 segments: canonical order = write order
    if nrows > (B * S + 1) * 10:
        pytest.skip("canonical order differs from write order")
    with O.Oracle(tmp_path) as orc:
        for preds, mask in [([], np.ones(nrows, bool)), ([("age", O.OP_GT, 18.0)], cols["age"] > 18), ([("age", O.OP_EQ, 127.0), ("age", O.OP_LT, 0.0)], np.zeros(nrows, bool)),
                            ([("state", O.OP_MATCH, ["CA", "NY"]), ("id", O.OP_LT, 0.0)], np.isin(cols["state"], [b"CA", b"NY"]) & (cols["id"] < 0))]:
            for aggs, group_by in [([(O.AGG_MIN, "age"), (O.AGG_MAX, "age")], ["state"]), ([(O.AGG_COUNT, "state")], []), ([(O.AGG_MAX, "id"), (O.AGG_COUNT, "id"), (O.AGG_MIN, "id")], ["age"]),
                                   ([(O.AGG_COUNT, "age"), (O.AGG_MIN, "id")], ["state", "age"])]:
                got = orc.query_agg("t", preds, aggs, group_by)
                want = _numpy_agg(cols, mask, aggs, group_by)
                assert got.nrows == len(want)
                for i, (k, vals) in enumerate(want):
                    for g in range(len(group_by)):
                        assert got.columns[g][i] == k[g]
                    for a, v in enumerate(vals):
                        assert got.columns[len(group_by) + a][i] == v
                for a, (op, _) in enumerate(aggs):
                    assert got.types[len(group_by) + a] == (O.COL_COUNT if op == O.AGG_COUNT else O.COL_DOUBLE)


def test_oracle_aggregation_rejects_what_the_reference_cannot_resolve(tmp_path):
    make_table(tmp_path, "t", 100, 8, 2, seed=1)
    with O.Oracle(tmp_path) as orc:
        with pytest.raises(O.OracleError):
            orc.query_agg("t", [], [(3, "age")], [])            # Sum: "Unknown Aggregate type" (Engine.scala:153)
        with pytest.raises(O.OracleError):
            orc.query_agg("t", [], [(O.AGG_MIN, "state")], [])   # resolved to MaxStringAggr in the reference (Engine.scala:147)
        with pytest.raises(O.OracleError):
            orc.query_agg("t", [], [(O.AGG_MIN, "nope")], [])
