"""Real OR (imm3_query_begin_dnf) - an extension: the reference drops the And / Or tag (Engine.scala:236-245), so the
parity path evaluates Or as And (tests/test_gpu_parity.py: `Or(...)` -> empty result).  Here the disjunction itself is checked:
expected rows = the oracle's unfiltered rows under the OR of the oracle's own per-conjunction selection bitmaps
(orc_filter_bitmap restates SelectOp per conjunction), cut at LIMIT in canonical order."""
import numpy as np
import pytest

import oracle_lib as O
from helpers import conj, make_table, oracle_preds
from immutable3_b200 import (And, EQ, Engine, GT, Imm3Error, LT, Match, NoSelect, Or, Project, Query, SegmentManager, Select, select_dnf)
from immutable3_b200 import _lib as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world(tmp_path_factory):
    d = tmp_path_factory.mktemp("real_or")
    rng = np.random.default_rng(17)
    n = 70_001
    extra = [("name:DENSE_STRING:size=5", np.array([b"alice", b"bobby", b"carol", b"dave_", b"erin#"], "S5")[rng.integers(0, 5, n)]),
             ("score:DENSE_INT", rng.integers(-1000, 1000, n).astype(np.int32))]
    make_table(d, "t", n, 100, 7, seed=8, extra_cols=extra)                                  # dense table, 100 segments, ragged blocks
    make_table(d, "p", 40_000, 1024, 3, seed=4, id_codec="PFOR_INT", id_mode="steps")        # sorted-int codec id: row-space filter + block emit
    with O.Oracle(d) as orc, SegmentManager(d) as sm:
        yield d, orc, sm


def expected(orc, table, select, proj, limit):
    n = orc.nrows(table)
    keep = np.zeros(n, dtype=bool)
    for term in select_dnf(select):
        words, _ = orc.filter_bitmap(table, oracle_preds(conj(*term)))
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")[:n].astype(bool)
        keep |= bits
    rows = orc.query(table, [], proj)
    cols = [c[keep] for c in rows.columns]
    return [c[:limit] for c in cols] if limit > 0 else cols


QUERIES = [
    ("t", Or(Select("age", GT(90)), Select("age", LT(5))), ["id", "age"]),
    ("t", Or(Select("state", Match(["CA"])), Select("state", Match(["DC", "VA"]))), ["id", "state"]),
    ("t", Or(And(Select("age", GT(18)), Select("age", LT(30))), And(Select("state", Match(["CA"])), Select("age", GT(60)))), ["id", "state", "age"]),
    ("t", And(Or(Select("age", LT(3)), Select("score", GT(900))), Or(Select("name", Match(["alice"])), Select("name", Match(["erin#"])))), ["id", "name", "score", "age"]),
    ("t", Or(Select("id", LT(50)), Or(Select("id", GT(69_000 * 3)), Select("score", EQ(7)))), ["id", "score"]),
    ("t", Or(Select("age", GT(200)), Select("age", EQ(7))), ["id"]),                            # first term can never hold (narrowed away)
    ("t", Or(Select("state", Match(["CAL"])), Select("state", Match(["X"]))), ["id"]),          # no term can hold: empty
    ("t", Or(Select("age", EQ(7)), NoSelect), ["age"]),                                         # `... or true`: every row
    ("t", Or(Select("age", GT(50)), Select("age", GT(50))), ["age"]),                           # the same term twice
    ("p", Or(Select("age", LT(2)), Select("state", Match(["CA"]))), ["id", "age", "state"]),
    ("p", Or(And(Select("age", GT(97)), Select("state", Match(["TX", "NY"]))), Select("age", EQ(0))), ["id"]),
]


@pytest.mark.parametrize("qi", range(len(QUERIES)))
def test_disjunctions_match_the_union_of_the_oracle_bitmaps(world, qi):
    d, orc, sm = world
    table, sel, proj = QUERIES[qi]
    eng = Engine(sm)
    for limit in (0, 1, 77, 10_000):
        exp = expected(orc, table, sel, proj, limit)
        with eng.execute(Query(table, sel, Project(proj, limit)), real_or=True) as got:
            assert got.nrows == len(exp[0]), (table, sel, limit, got.nrows, len(exp[0]))
            for c in range(len(proj)):
                assert np.array_equal(got.column(c), exp[c]), (table, sel, proj[c], limit)
    # the parity path is untouched: Or == And there (an empty result for disjoint ranges)
    if qi == 0:
        with eng.execute(Query(table, sel, Project(proj))) as got:
            assert got.nrows == 0


def test_a_disjunction_over_an_encoded_column_is_refused(world):
    d, orc, sm = world
    with pytest.raises(Imm3Error) as e:
        Engine(sm).execute(Query("p", Or(Select("id", LT(100)), Select("age", EQ(3))), Project(["id"])), real_or=True)
    assert e.value.status == L.ERR_UNSUPPORTED
    # a single conjunction through the same entry point takes every kernel path of the library
    exp = orc.query("p", oracle_preds(conj(Select("id", GT(2000)), Select("id", LT(30_000)))), ["id"])
    with Engine(sm).execute(Query("p", conj(Select("id", GT(2000)), Select("id", LT(30_000))), Project(["id"])), real_or=True) as got:
        assert got.nrows == exp.nrows and np.array_equal(got.column(0), exp.columns[0])


def test_full_size_disjunction(tmp_path_factory):
    """40 M rows (4883 tiles: the offset scan is a kernel of its own), three terms, against numpy on the files."""
    from oracle_lib import synth_write

    d = tmp_path_factory.mktemp("or40m")
    n = 40_000_000
    synth_write(d, "syn", n)
    order = sorted(range(40), key=lambda i: f"id_{i}.dat")
    age = np.concatenate([np.fromfile(d / "syn" / f"age_{i}.dat", np.int8) for i in order])
    ids = np.concatenate([np.fromfile(d / "syn" / f"id_{i}.dat", "<i4") for i in order])
    st = np.concatenate([np.fromfile(d / "syn" / f"state_{i}.dat", "S2") for i in order])
    sel = Or(And(Select("age", GT(18)), Select("age", LT(30))), Or(And(Select("state", Match(["CA"])), Select("age", GT(90))), Select("id", LT(1000))))
    keep = ((age > 18) & (age < 30)) | ((st == b"CA") & (age > 90)) | (ids < 1000)
    with SegmentManager(d) as sm:
        for limit in (0, 3_000_000):
            with Engine(sm).execute(Query("syn", sel, Project(["id", "age"], limit)), real_or=True) as got:
                want = ids[keep][:limit] if limit else ids[keep]
                assert got.nrows == len(want) and np.array_equal(got.column(0), want)
                assert np.array_equal(got.column(1), age[keep][:limit] if limit else age[keep])
