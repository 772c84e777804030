"""Oracle query semantics against the quirks catalogued in SURVEY.md §3.4 and an independent numpy
evaluation — this is what 'the reference's results' means for the GPU parity tests."""
import numpy as np
import pytest

import oracle_lib as O
from helpers import conj, make_table, numpy_expected, oracle_preds
from immutable3_b200 import And, EQ, GT, LT, Match, NoSelect, Or, Select


@pytest.fixture(scope="module")
def small(tmp_path_factory):
    d = tmp_path_factory.mktemp("sem")
    cols = make_table(d, "t", 3000, 64, 5, seed=2)          # 10 segments: canonical order == write order
    cols_neg = make_table(d, "neg", 2500, 32, 10, seed=3, id_mode="random")  # ages over the full int8 range
    return d, cols, cols_neg


def test_readme_query_config1(tmp_path):
    # README.md:6 on a one-block test_100: in the reference's well-defined domain (ref_throw == 0)
    cols = make_table(tmp_path, "test_100", 100, 1024, 1000, seed=2)   # 11 matching rows: LIMIT 10 is reached
    with O.Oracle(tmp_path) as orc:
        r = orc.query("test_100", [("age", O.OP_GT, 18), ("age", O.OP_LT, 30)], ["id", "age"], limit=10, fmt_rows=10)
    (eid, eage), _ = numpy_expected(cols, conj(Select("age", GT(18)), Select("age", LT(30))), ["id", "age"], 10)
    assert np.array_equal(r.columns[0], eid) and np.array_equal(r.columns[1], eage)
    assert r.ref_throw == 0 and r.ref_rows == 10
    assert r.format_rows() == [f"Row({i},{a})" for i, a in zip(eid, eage)]  # Record.scala:13
    # with only 9 matches in the block the reference runs off the end of segment 0 and throws (B1)
    cols = make_table(tmp_path, "test_100b", 100, 1024, 1000, seed=4)
    with O.Oracle(tmp_path) as orc:
        r = orc.query("test_100b", [("age", O.OP_GT, 18), ("age", O.OP_LT, 30)], ["id", "age"], limit=10)
    assert r.nrows == 9 and r.ref_throw == 1 and r.ref_rows == 9


def _simulate_reference(sel_mask_seg0, block_rows, limit):
    """ProjectIterator over ResultQueueOp with --cpu-count 1 (Project.scala:37-80, ResultQueue.scala:15-29)."""
    emitted, at = 0, 0
    for n in block_rows:
        sel = int(sel_mask_seg0[at:at + n].sum())
        at += n
        if sel == 0:
            return 2, emitted           # B2: empty batch -> ArrayIndexOutOfBounds
        emitted += sel if limit <= 0 else min(sel, limit - emitted)
        if limit > 0 and emitted >= limit:
            return 0, emitted           # LIMIT reached: clean stop
    return 1, emitted                   # B1: None.get at the end-of-segment marker


def test_reference_would_throw_reports(small):
    d, cols, _ = small
    blocks = [64] * 5 + [1]  # segment 0: S full blocks + the 1-row tail block
    n0 = sum(blocks)
    with O.Oracle(d) as orc:
        for preds, mask in (([("age", O.OP_GT, 50)], cols["age"] > 50), ([("age", O.OP_GT, -1)], cols["age"] > -1),
                            ([("age", O.OP_EQ, 127)], cols["age"] == 127), ([], np.ones(len(cols["age"]), bool))):
            for limit in (0, 5, 100, 100000):
                r = orc.query("t", preds, ["id"], limit=limit)
                assert (r.ref_throw, r.ref_rows) == _simulate_reference(mask[:n0], blocks, limit), (preds, limit)
        # no predicate, no LIMIT: every row of segment 0 is printed, then B1
        r = orc.query("t", [], ["id"])
        assert r.ref_throw == 1 and r.ref_rows == n0 and r.nrows == len(cols["id"])


def test_predicates_match_numpy(small):
    d, cols, cols_neg = small
    queries = [
        conj(Select("age", GT(18)), Select("age", LT(30))),
        conj(Select("state", Match(["CA"])), Select("age", GT(18)), Select("age", LT(30))),
        conj(Select("state", Match(["DC", "CT"])), Select("age", GT(0))),     # Engine.scala:39-46
        conj(Select("id", GT(1000)), Select("id", LT(4000))),
        conj(Select("age", EQ(7))),
        conj(Select("id", EQ(5 + 3 * 1234))),
        NoSelect,
    ]
    with O.Oracle(d) as orc:
        for q in queries:
            for limit in (0, 1, 10, 100000):
                r = orc.query("t", oracle_preds(q), ["id", "state", "age"], limit=limit)
                exp, nm = numpy_expected(cols, q, ["id", "state", "age"], limit)
                assert all(np.array_equal(a, b) for a, b in zip(r.columns, exp)), (q, limit)
                if limit == 0:
                    assert r.nmatched == nm
        for q in [conj(Select("age", GT(200))), conj(Select("age", EQ(300))), conj(Select("age", LT(1e10))),
                  conj(Select("id", GT(3e9))), conj(Select("id", LT(3e9))), conj(Select("age", GT(-129))),
                  conj(Select("id", GT(-3e9)), Select("age", LT(-5)))]:
            r = orc.query("neg", oracle_preds(q), ["id", "age"])
            exp, _ = numpy_expected(cols_neg, q, ["id", "age"])
            assert all(np.array_equal(a, b) for a, b in zip(r.columns, exp)), q


def test_narrowing_known_answers(small):
    d, _, cols = small
    with O.Oracle(d) as orc:
        # GT(200.0) on TINYINT compares against -56; EQ(300.0) matches 44; LT(1e10) -> threshold -1
        assert orc.query("neg", [("age", O.OP_GT, 200.0)], ["age"]).nrows == int((cols["age"] > -56).sum())
        assert orc.query("neg", [("age", O.OP_EQ, 300.0)], ["age"]).nrows == int((cols["age"] == 44).sum())
        assert orc.query("neg", [("age", O.OP_LT, 1e10)], ["age"]).nrows == int((cols["age"] < -1).sum())
        # GT(3e9) on INT: threshold 2147483647 matches nothing; LT(3e9) everything but 2147483647
        assert orc.query("neg", [("id", O.OP_GT, 3e9)], ["id"]).nrows == 0
        assert orc.query("neg", [("id", O.OP_LT, 3e9)], ["id"]).nrows == int((cols["id"] != 2**31 - 1).sum())


def test_match_length_and_or_is_and(small):
    d, cols, _ = small
    with O.Oracle(d) as orc:
        assert orc.query("t", [("state", O.OP_MATCH, ["CAL"])], ["id"]).nrows == 0           # Select.scala:37
        assert orc.query("t", [("state", O.OP_MATCH, ["C"])], ["id"]).nrows == 0
        q = Or(Select("age", GT(50)), Select("age", LT(10)))                                  # Engine.scala:240
        assert orc.query("t", oracle_preds(q), ["id"]).nrows == 0
        q2 = Or(Select("age", GT(10)), Select("age", LT(50)))
        assert orc.query("t", oracle_preds(q2), ["id"]).nrows == int(((cols["age"] > 10) & (cols["age"] < 50)).sum())


def test_type_errors_and_unknown_names(small):
    d, _, _ = small
    with O.Oracle(d) as orc:
        for preds in ([("state", O.OP_GT, 1)], [("state", O.OP_EQ, 1)], [("age", O.OP_MATCH, ["x"])], [("id", O.OP_MATCH, ["x"])]):
            with pytest.raises(O.OracleError) as e:
                orc.query("t", preds, ["id"])
            assert e.value.status == -2 and "Unsupported column vector" in str(e.value)
        with pytest.raises(O.OracleError) as e:
            orc.query("t", [("state", O.OP_NOTMATCH, ["CA"])], ["id"])
        assert "Unsupported condition" in str(e.value)
        with pytest.raises(O.OracleError) as e:
            orc.query("t", [], ["nope"])
        assert e.value.status == -1 and "Column nope does not exist in table t" in str(e.value)
        with pytest.raises(O.OracleError) as e:
            orc.query("missing", [], ["id"])
        assert "Table missing does not exist in SegmentManager" in str(e.value)


def test_segment_order_is_lexicographic(tmp_path):
    # 12 segments: 0,1,10,11,2,...,9 (SegmentManager.scala:41)
    n = 12 * (8 * 2 + 1) - 3
    cols = make_table(tmp_path, "t", n, 8, 2)
    with O.Oracle(tmp_path) as orc:
        assert orc.segment_file_ids("t") == [0, 1, 10, 11, 2, 3, 4, 5, 6, 7, 8, 9]
        r = orc.query("t", [], ["id"])
        per = 17
        want = np.concatenate([cols["id"][s * per:(s + 1) * per] for s in [0, 1, 10, 11, 2, 3, 4, 5, 6, 7, 8, 9]])
        assert np.array_equal(r.columns[0], want)
        # LIMIT cuts in canonical order; threads do not change the answer
        for th in (1, 3, 8):
            r2 = orc.query("t", [("age", O.OP_GT, 20)], ["id", "age"], limit=40, nthreads=th)
            m = cols["age"] > 20
            canon = np.concatenate([np.arange(s * per, min(n, (s + 1) * per)) for s in [0, 1, 10, 11, 2, 3, 4, 5, 6, 7, 8, 9]])
            idx = canon[m[canon]][:40]
            assert np.array_equal(r2.columns[0], cols["id"][idx]) and np.array_equal(r2.columns[1], cols["age"][idx])
        # shard slices concatenate to the whole
        parts = [orc.query("t", [("age", O.OP_GT, 20)], ["id"], seg_begin=a, seg_end=b).columns[0] for a, b in ((0, 5), (5, 12))]
        assert np.array_equal(np.concatenate(parts), orc.query("t", [("age", O.OP_GT, 20)], ["id"]).columns[0])


def test_projection_order_duplicates_and_bitmap(small):
    d, cols, _ = small
    with O.Oracle(d) as orc:
        r = orc.query("t", [("age", O.OP_LT, 5)], ["age", "id", "age"])
        m = cols["age"] < 5
        assert np.array_equal(r.columns[0], cols["age"][m]) and np.array_equal(r.columns[1], cols["id"][m]) and np.array_equal(r.columns[2], cols["age"][m])
        words, nsel = orc.filter_bitmap("t", [("age", O.OP_LT, 5)])
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")[: len(m)].astype(bool)
        assert np.array_equal(bits, m) and nsel == m.sum()
