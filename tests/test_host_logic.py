"""Host-side logic of libimm3gpu.so that needs no GPU: metadata parsing, canonical segment order,
sharding, validation, the planner (narrowing + predicate merging) and the SQL grammar — opened with
IMM3_OPEN_HOST_ONLY (no device work; queries refuse to run: there is no CPU fallback)."""
import json
import os
import shutil

import numpy as np
import pytest

import oracle_lib as O
from helpers import conj, make_table
from immutable3_b200 import (And, EQ, Engine, GT, Imm3Error, LT, Match, NoSelect, NotMatch, OPEN_HOST_ONLY, Or, Project, Query,
                             SegmentManager, Select)
from immutable3_b200 import _lib as L
from immutable3_b200.dist import limit_split, shard_range


@pytest.fixture(scope="module")
def db(tmp_path_factory):
    d = tmp_path_factory.mktemp("host")
    n = 12 * (8 * 2 + 1) - 3
    make_table(d, "t", n, 8, 2)                                   # 12 segments
    make_table(d, "p", 1000, 64, 3, id_codec="PFOR_INT", id_mode="steps")
    sm = SegmentManager(d, flags=OPEN_HOST_ONLY)
    yield d, sm, n
    sm.close()


def test_tables_columns_and_counts(db):
    d, sm, n = db
    assert sorted(t.name for t in sm.tables) == ["p", "t"]
    t = sm.getTable("t")
    assert [(c.name, c.columnType, c.codec, c.width) for c in t.columns] == [
        ("id", "INT", "DENSE_INT", 4), ("state", "STRING", "DENSE_STRING", 2), ("age", "TINYINT", "DENSE_TINYINT", 1)]
    assert (t.blockSize, t.nsegments, t.seg_begin, t.seg_end, t.nrows) == (8, 12, 0, 12, n)
    assert t.nblocks == 11 * 3 + 2 and t.resident_bytes == 7 * n
    assert sm.getTableSegmentCount("t") == 12
    p = sm.getTable("p")
    assert p.columns[0].codec == "PFOR_INT" and p.nrows == 1000
    with O.Oracle(d) as orc:
        assert orc.nrows("t") == n and orc.nrows("p") == 1000 and orc.nsegments("t") == 12
    with pytest.raises(Imm3Error) as e:
        sm.getTable("missing")
    assert e.value.status == L.ERR_NOT_FOUND and "Table missing does not exist in SegmentManager" in e.value.message


def test_canonical_segment_order_is_lexicographic(db):
    d, sm, _ = db
    assert sm.segmentFileIds("t") == [0, 1, 10, 11, 2, 3, 4, 5, 6, 7, 8, 9]   # SegmentManager.scala:41
    with O.Oracle(d) as orc:
        assert orc.segment_file_ids("t") == sm.segmentFileIds("t")


def test_shard_slices_partition_the_canonical_list(db):
    d, _, n = db
    with O.Oracle(d) as orc:
        for world in (1, 2, 3, 5, 8, 12, 16):
            rows, prev_end = 0, 0
            for rank in range(world):
                with SegmentManager(d, rank=rank, world=world, flags=OPEN_HOST_ONLY) as s:
                    t = s.getTable("t")
                    assert (t.seg_begin, t.seg_end) == shard_range(12, rank, world) and t.seg_begin == prev_end
                    assert t.nrows == orc.nrows("t", t.seg_begin, t.seg_end)
                    rows += t.nrows
                    prev_end = t.seg_end
            assert rows == n and prev_end == 12


def test_limit_split_arithmetic():
    assert limit_split([5, 0, 7], 0) == ([0, 5, 5], [5, 0, 7])
    assert limit_split([5, 0, 7], 6) == ([0, 5, 5], [5, 0, 1])
    assert limit_split([5, 5, 5], 5) == ([0, 5, 10], [5, 0, 0])
    assert limit_split([0, 0, 3], 10) == ([0, 0, 0], [0, 0, 3])


def _explain(sm, table, select, cols=("id",), limit=0):
    return Engine(sm).explain(Query(table, select, Project(list(cols), limit)))


def test_planner_narrows_like_the_jvm(db):
    _, sm, _ = db
    # GT(200.0)/EQ(300.0)/LT(1e10) on TINYINT -> -56 / 44 / -1; GT(3e9)/LT(3e9) on INT -> 2147483647
    ex = _explain(sm, "t", conj(Select("age", GT(200.0)), Select("age", EQ(300.0)), Select("age", LT(1e10))))
    assert [n["byte"] for n in ex["narrowed"]] == [-56, 44, -1]
    ex = _explain(sm, "t", conj(Select("id", GT(3e9)), Select("id", LT(-3e9)), Select("id", EQ(float("nan"))), Select("id", GT(18.9)), Select("id", LT(-18.9))))
    assert [n["int"] for n in ex["narrowed"]] == [2147483647, -2147483648, 0, 18, -18]
    for d, want in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat_vectors.json")))["d2b"]:
        assert _explain(sm, "t", Select("age", EQ(d)))["narrowed"][0]["byte"] == want == O.lib().orc_d2b(d)


def test_planner_merges_a_conjunction_into_one_range_per_column(db):
    _, sm, _ = db
    ex = _explain(sm, "t", conj(Select("age", GT(18)), Select("age", LT(30))), ("id", "age"), 10)
    assert ex["filters"] == [{"col": "age", "kind": "i8_range", "lo": 19, "hi": 29}] and ex["limit"] == 10 and not ex["always_empty"]
    assert ex["kernel"].startswith("filter(") and ex["proj"] == ["id", "age"]
    ex = _explain(sm, "t", Or(Select("age", GT(50)), Select("age", LT(10))))         # OR == AND (Engine.scala:240)
    assert ex["filters"][0]["lo"] == 51 and ex["filters"][0]["hi"] == 9 and ex["always_empty"]
    ex = _explain(sm, "t", conj(Select("id", GT(3e9))))
    assert ex["always_empty"] and ex["kernel"].startswith("none")
    ex = _explain(sm, "t", conj(Select("id", LT(3e9))))
    assert ex["filters"][0] == {"col": "id", "kind": "i32_range", "lo": -2**31, "hi": 2**31 - 2}
    ex = _explain(sm, "t", conj(Select("state", Match(["CA", "CAL", "NY", "CA"])), Select("age", EQ(7)), Select("state", Match(["NY", "CA", "TX"]))))
    assert ex["filters"] == [{"col": "state", "kind": "str_match", "lits": ["CA", "NY"]}, {"col": "age", "kind": "i8_range", "lo": 7, "hi": 7}]
    assert _explain(sm, "t", Select("state", Match(["CAL"])))["always_empty"]         # Select.scala:37
    assert _explain(sm, "p", Select("id", GT(5)))["kernel"].startswith("blocks_filter")  # sorted-int codec -> block pipeline
    assert _explain(sm, "t", NoSelect)["filters"] == []


def test_planner_errors_follow_the_reference(db):
    _, sm, _ = db
    for sel in (Select("state", GT(1)), Select("state", EQ(1)), Select("age", Match(["x"])), Select("id", Match(["x"]))):
        with pytest.raises(Imm3Error) as e:
            _explain(sm, "t", sel)
        assert e.value.status == L.ERR_UNSUPPORTED and "Unsupported column vector" in e.value.message   # Select.scala:41,80,118,156
    with pytest.raises(Imm3Error) as e:
        _explain(sm, "t", Select("state", NotMatch(["CA"])))
    assert e.value.status == L.ERR_UNSUPPORTED and "Unsupported condition" in e.value.message              # Select.scala:22
    with pytest.raises(Imm3Error) as e:
        _explain(sm, "t", NoSelect, ("nope",))
    assert e.value.status == L.ERR_NOT_FOUND and "Column nope does not exist in table t" in e.value.message  # Table.scala:13
    with pytest.raises(Imm3Error) as e:
        _explain(sm, "t", Select("nope", GT(1)))
    assert e.value.status == L.ERR_NOT_FOUND


def test_no_cpu_fallback_on_host_only_handle(db):
    _, sm, _ = db
    with pytest.raises(Imm3Error) as e:
        Engine(sm).execute(Query("t", NoSelect, Project(["id"])))
    assert e.value.status == L.ERR_STATE and "no CPU fallback" in e.value.message
    with pytest.raises(Imm3Error):
        Engine(sm).filter_bitmap("t", NoSelect)


def _bad_copy(src, dst):
    shutil.copytree(src, dst)
    return dst


def test_validation_rejects_what_the_reference_leaves_undefined(db, tmp_path):
    d, _, _ = db
    def expect_bad(mutate, needle):
        root = tmp_path / needle.replace(" ", "_")[:20]
        _bad_copy(d, root)
        shutil.rmtree(root / "p")
        mutate(root / "t")
        with pytest.raises(Imm3Error) as e:
            SegmentManager(root, flags=OPEN_HOST_ONLY)
        assert e.value.status in (L.ERR_BAD_FORMAT, L.ERR_IO) and needle in e.value.message, e.value.message

    expect_bad(lambda t: os.remove(t / "age_3.meta"), ".meta files")
    expect_bad(lambda t: (os.remove(t / "age_3.meta"), os.remove(t / "age_3.dat")), "segments")
    expect_bad(lambda t: open(t / "id_2.meta", "w").write('{"blockOffset":[0,32,64,67]}'), "multiple of the value width")
    expect_bad(lambda t: open(t / "id_2.meta", "w").write('{"blockOffset":[4,32,64,68]}'), "start at 0")
    expect_bad(lambda t: open(t / "id_2.meta", "w").write('{"blockOffset":[0,64,32,68]}'), "decreases")
    expect_bad(lambda t: open(t / "id_2.meta", "w").write('{"blockOffset":[0,32,64,68,72]}'), "block offsets need")
    expect_bad(lambda t: open(t / "age_2.meta", "w").write('{"blockOffset":[0,9,16,17]}'), "same rows per block")
    expect_bad(lambda t: open(t / "_table.meta", "w").write("{not json"), "invalid JSON")
    expect_bad(lambda t: open(t / "_table.meta", "w").write('{"name":"t","columns":[{"name":"id","columnType":"INT","codec":"SNAPPY","dtypeAttrs":{}}],"blockSize":8}'), "unknown codec")
    expect_bad(lambda t: open(t / "_table.meta", "w").write('{"name":"t","columns":[{"name":"state","columnType":"STRING","codec":"DENSE_STRING","dtypeAttrs":{}}],"blockSize":8}'), "size")


def test_validation_of_pfor_blocks(db, tmp_path):
    d, _, _ = db
    root = tmp_path / "pf"
    shutil.copytree(d, root)
    shutil.rmtree(root / "t")
    raw = bytearray(open(root / "p" / "id_0.dat", "rb").read())
    raw[7] = 99  # first header word: a bit width of 99
    open(root / "p" / "id_0.dat", "wb").write(raw)
    with pytest.raises(Imm3Error) as e:
        SegmentManager(root, flags=OPEN_HOST_ONLY)
    assert e.value.status == L.ERR_BAD_FORMAT and "PFOR_INT" in e.value.message


def test_table_meta_accepts_whole_doubles_and_ints(tmp_path):
    make_table(tmp_path, "t", 10, 4, 2)
    p = tmp_path / "t" / "_table.meta"
    p.write_text(p.read_text().replace('"blockSize":4', '"blockSize":4.0'))
    for f in os.listdir(tmp_path / "t"):
        if f.endswith(".meta") and f != "_table.meta":
            q = tmp_path / "t" / f
            q.write_text(q.read_text().replace(",", ".0, ").replace("]", ".0 ]"))
    with SegmentManager(tmp_path, flags=OPEN_HOST_ONLY) as sm:
        assert sm.getTable("t").nrows == 10 and sm.getTable("t").blockSize == 4


# ---- SQL grammar (SQLParser.scala) through imm3_query_sql: parse errors surface before any device work ----
def _sql_status(sm, sql):
    try:
        Engine(sm).execute_sql(sql)
    except Imm3Error as e:
        return e.status, e.message
    return 0, ""


def test_sql_grammar(db):
    _, sm, _ = db
    ok = L.ERR_STATE  # parsed + planned fine, then refused because the handle is host-only
    assert _sql_status(sm, "select id, age from t where (age > 18 and age < 30) limit 10")[0] == ok      # README.md:6
    assert _sql_status(sm, "select id,state,age from t where (state = 'CA' and age > 18 and age < 30)")[0] == ok
    assert _sql_status(sm, "select id from t")[0] == ok
    assert _sql_status(sm, "select id from t where age = 7")[0] == ok
    assert _sql_status(sm, "select id from t where (age > 1 and (age < 9 or age = 3))")[0] == ok
    assert _sql_status(sm, "  select id from t where age > 1e1 limit 3  ")[0] == ok
    assert _sql_status(sm, "select id from t where age > 10d")[0] == ok                                 # Double.parseDouble suffix
    st, msg = _sql_status(sm, "select id from t where state = CA")
    assert st == L.ERR_INVALID_ARG and "NumberFormatException" in msg                                    # "CA".toDouble
    assert _sql_status(sm, "select id from t where age >= 3")[0] == L.ERR_INVALID_ARG                    # no >= in the grammar
    assert _sql_status(sm, "select id from t where age > -3")[0] == L.ERR_INVALID_ARG                    # no minus sign
    assert _sql_status(sm, "select id from t where age > 3.5")[0] == L.ERR_INVALID_ARG                   # no decimal point
    assert _sql_status(sm, "SELECT id FROM t")[0] == L.ERR_INVALID_ARG                                   # lowercase keywords only
    assert _sql_status(sm, "select id from t where () limit 1")[0] == L.ERR_INVALID_ARG
    assert _sql_status(sm, "select id from t limit 99999999999")[0] == L.ERR_INVALID_ARG
    assert _sql_status(sm, "select min(age) from t")[0] == L.ERR_UNSUPPORTED                             # ProjectAgg: out of scope
    assert _sql_status(sm, "select id from nope")[0] == L.ERR_NOT_FOUND
    assert _sql_status(sm, "select id from t where state > 3")[0] == L.ERR_UNSUPPORTED


def test_select_tree_in_disjunctive_normal_form():
    """select_dnf (real OR, imm3_query_begin_dnf): And distributes over Or, leaves keep their application order, NoSelect is `true`."""
    from immutable3_b200 import select_dnf

    a, b, c, d_ = Select("age", GT(1)), Select("age", LT(9)), Select("state", Match(["CA"])), Select("id", EQ(4))
    assert select_dnf(NoSelect) == [[]]
    assert select_dnf(a) == [[a]]
    assert select_dnf(And(a, b)) == [[a, b]]
    assert select_dnf(Or(a, b)) == [[a], [b]]
    assert select_dnf(And(Or(a, b), c)) == [[a, c], [b, c]]
    assert select_dnf(And(Or(a, b), Or(c, d_))) == [[a, c], [a, d_], [b, c], [b, d_]]
    assert select_dnf(Or(And(a, b), And(c, Or(d_, a)))) == [[a, b], [c, d_], [c, a]]
    assert select_dnf(Or(a, NoSelect)) == [[a], []]
    big = a
    for _ in range(5):
        big = And(Or(big, b), Or(c, d_))
    with pytest.raises(Imm3Error) as e:
        select_dnf(big)
    assert e.value.status == L.ERR_UNSUPPORTED
