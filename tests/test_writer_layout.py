"""Writer side (SegmentWriter + LoaderCli roll logic + PFORCodecInt.encode) against the layout
known-answers of SURVEY.md §3.5/§8c, against an independent numpy restatement written here, and
against the oracle reading the files back."""
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle_lib as O
from helpers import make_table
from immutable3_b200 import _lib as L
from immutable3_b200.loader import SegmentWriter, load_csv, pfor_encode, synth_rows, synth_segments, synth_write


def _meta(path):
    return json.load(open(path))["blockOffset"]


def test_loader_layout_known_answer(tmp_path):
    # B=4, S=2, 20 rows: Segment.scala:99-151 + LoaderCli.scala:142-148
    with SegmentWriter(tmp_path, "t", ["a:DENSE_INT"], 4, 2) as w:
        w.append(np.arange(20, dtype=np.int32))
    d = tmp_path / "t"
    assert sorted(os.listdir(d)) == ["_table.meta", "a_0.dat", "a_0.meta", "a_1.dat", "a_1.meta", "a_2.dat", "a_2.meta"]
    assert _meta(d / "a_0.meta") == [0, 16, 32, 36]
    assert _meta(d / "a_1.meta") == [0, 16, 32, 36]
    assert _meta(d / "a_2.meta") == [0, 8]
    assert np.array_equal(np.fromfile(d / "a_0.dat", "<i4"), np.arange(0, 9))
    assert np.array_equal(np.fromfile(d / "a_1.dat", "<i4"), np.arange(9, 18))
    assert np.array_equal(np.fromfile(d / "a_2.dat", "<i4"), np.arange(18, 20))
    tm = json.load(open(d / "_table.meta"))
    assert tm == {"name": "t", "columns": [{"name": "a", "columnType": "INT", "codec": "DENSE_INT", "dtypeAttrs": {}}], "blockSize": 4}


def test_table_meta_text_is_ujson_compact(tmp_path):
    with SegmentWriter(tmp_path, "test_100m", ["id:DENSE_INT", "state:DENSE_STRING:size=2", "age:DENSE_TINYINT"], 1024, 1000):
        pass
    txt = open(tmp_path / "test_100m" / "_table.meta").read()
    assert txt == ('{"name":"test_100m","columns":[{"name":"id","columnType":"INT","codec":"DENSE_INT","dtypeAttrs":{}},'
                   '{"name":"state","columnType":"STRING","codec":"DENSE_STRING","dtypeAttrs":{"size":"2"}},'
                   '{"name":"age","columnType":"TINYINT","codec":"DENSE_TINYINT","dtypeAttrs":{}}],"blockSize":1024}')


def _numpy_layout(nrows, B, S):
    """Independent restatement: every full segment = S full blocks + a 1-row tail block."""
    per = S * B + 1
    segs = []
    r = 0
    while r < nrows:
        n = min(per, nrows - r)
        blocks = [B] * (n // B) + ([n % B] if n % B else [])
        segs.append((r, n, blocks))
        r += n
    return segs


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 7), st.integers(1, 5), st.integers(0, 200), st.integers(1, 50))
def test_layout_matches_independent_restatement(tmp_path_factory, B, S, nrows, chunk):
    d = tmp_path_factory.mktemp("lay")
    ids = np.arange(nrows, dtype=np.int32) * 7 - 3
    ages = (np.arange(nrows) % 251 - 125).astype(np.int8)
    with SegmentWriter(d, "t", ["id:DENSE_INT", "age:DENSE_TINYINT"], B, S) as w:
        for a in range(0, nrows, chunk):  # chunking must not matter
            w.append(ids[a:a + chunk], ages[a:a + chunk])
    segs = _numpy_layout(nrows, B, S) or [(0, 0, [])]  # an empty table still has an empty segment 0
    for i, (r0, n, blocks) in enumerate(segs):
        assert _meta(d / "t" / f"id_{i}.meta") == [0] + list(np.cumsum([4 * b for b in blocks]))
        assert _meta(d / "t" / f"age_{i}.meta") == [0] + list(np.cumsum(blocks))
        assert np.array_equal(np.fromfile(d / "t" / f"id_{i}.dat", "<i4"), ids[r0:r0 + n])
        assert np.array_equal(np.fromfile(d / "t" / f"age_{i}.dat", "i1"), ages[r0:r0 + n])
    assert not os.path.exists(d / "t" / f"id_{len(segs)}.dat")


def test_readme_layout_sizes(tmp_path):
    # README settings scaled down 16x in block size: S*B+1 rows per full segment, S+1 blocks, S+2 offsets
    B, S, n = 64, 1000, 64 * 1000 + 1 + 12345
    make_table(tmp_path, "t", n, B, S)
    assert os.path.getsize(tmp_path / "t" / "id_0.dat") == 4 * (B * S + 1)
    assert os.path.getsize(tmp_path / "t" / "state_0.dat") == 2 * (B * S + 1)
    assert os.path.getsize(tmp_path / "t" / "age_0.dat") == B * S + 1
    off = _meta(tmp_path / "t" / "age_0.meta")
    assert len(off) == S + 2 and off[-1] - off[-2] == 1
    assert os.path.getsize(tmp_path / "t" / "age_1.dat") == 12345


def test_csv_loader_discards_header_trims_and_matches_typed_append(tmp_path):
    rows = [(i, ["CA", "NY", "DC"][i % 3], (i * 7) % 100) for i in range(57)]
    csv = tmp_path / "in.csv"
    csv.write_text("id,state,age\n" + "\n".join(f" {a} ,{b}, {c}" for a, b, c in rows) + "\n")
    specs = ["id:DENSE_INT", "state:DENSE_STRING:size=2", "age:DENSE_TINYINT"]
    load_csv(tmp_path / "a", "t", specs, 8, 3, csv)
    with SegmentWriter(tmp_path / "b", "t", specs, 8, 3) as w:
        w.append(np.array([r[0] for r in rows], np.int32), np.array([r[1] for r in rows], "S2"), np.array([r[2] for r in rows], np.int8))
    for f in sorted(os.listdir(tmp_path / "a" / "t")):
        assert open(tmp_path / "a" / "t" / f, "rb").read() == open(tmp_path / "b" / "t" / f, "rb").read(), f


def test_csv_loader_rejects_what_the_reference_would_corrupt(tmp_path):
    csv = tmp_path / "in.csv"
    csv.write_text("h\nCAL\n")
    with pytest.raises(L.Imm3Error) as e:
        load_csv(tmp_path, "t", ["state:DENSE_STRING:size=2"], 8, 3, csv)
    assert e.value.status == L.ERR_INVALID_ARG
    csv.write_text("h\n300\n")
    with pytest.raises(L.Imm3Error):
        load_csv(tmp_path, "t", ["age:DENSE_TINYINT"], 8, 3, csv)


@settings(max_examples=150, deadline=None)
@given(st.lists(st.integers(-2**31, 2**31 - 1), min_size=0, max_size=400))
def test_product_pfor_encoder_equals_oracle_encoder(vals):
    # two independently written encoders (bit-stream accumulator vs per-word OR) must agree byte for byte
    assert pfor_encode(vals) == O.pfor_encode(vals)


def test_product_pfor_encoder_sorted_blocks():
    rng = np.random.default_rng(1)
    for n in (1, 31, 32, 33, 127, 128, 129, 160, 1023, 1024, 1025):
        v = np.cumsum(rng.integers(0, 1 << rng.integers(1, 20), size=n)).astype(np.int32)
        enc = pfor_encode(v)
        assert enc == O.pfor_encode(v)
        assert np.array_equal(O.pfor_decode(enc), v)


def test_pfor_column_files_decode_with_oracle(tmp_path):
    cols = make_table(tmp_path, "t", 3 * 64 * 5 + 77, 64, 5, id_codec="PFOR_INT", id_mode="steps")
    with O.Oracle(tmp_path) as orc:
        r = orc.query("t", [], ["id", "age", "state"])
    assert np.array_equal(r.columns[0], cols["id"]) and np.array_equal(r.columns[1], cols["age"]) and np.array_equal(r.columns[2], cols["state"])
    off = _meta(tmp_path / "t" / "id_0.meta")
    raw = open(tmp_path / "t" / "id_0.dat", "rb").read()
    assert np.array_equal(O.pfor_decode(raw[off[0]:off[1]]), cols["id"][:64])
    assert raw[off[1] - 8:off[1]] == b"\0" * 8  # every block ends with the 8 pad bytes of PFORCodec.scala:20


def test_synthetic_generator_is_counter_based_and_matches_files(tmp_path):
    n, B, S = 5000, 16, 10
    nseg = synth_segments(n, B, S)
    assert nseg == -(-n // (B * S + 1))
    one = tmp_path / "one"
    tmp_path = tmp_path / "coop"
    synth_write(tmp_path, "syn", n, B, S, write_table_meta=True, seg_id_begin=0, seg_id_end=0)  # meta only
    assert sorted(os.listdir(tmp_path / "syn")) == ["_table.meta"]
    # segments written out of order and by separate calls give the same files as one call
    for s in reversed(range(nseg)):
        synth_write(tmp_path, "syn", n, B, S, seg_id_begin=s, seg_id_end=s + 1, write_table_meta=False)
    synth_write(one, "syn", n, B, S)
    assert sorted(os.listdir(one / "syn")) == sorted(os.listdir(tmp_path / "syn"))
    for f in sorted(os.listdir(one / "syn")):
        assert open(tmp_path / "syn" / f, "rb").read() == open(one / "syn" / f, "rb").read(), f
    ids, ages, states = synth_rows(0, n)
    assert np.array_equal(ids, np.arange(n)) and ages.min() >= 0 and ages.max() < 100
    with O.Oracle(tmp_path) as orc:
        r = orc.query("syn", [], ["id", "age", "state"])
    # < 11 segments... no: canonical order is lexicographic, so compare as multisets keyed by id
    order = np.argsort(r.columns[0], kind="stable")
    assert np.array_equal(r.columns[0][order], ids) and np.array_equal(r.columns[1][order], ages) and np.array_equal(r.columns[2][order], states)
    # uniformity sanity: all 51 states, all 100 ages present
    assert len(set(states.tolist())) == 51 and len(set(ages.tolist())) == 100


def _dir_bytes(d):
    return {f: open(os.path.join(d, f), "rb").read() for f in sorted(os.listdir(d)) if not f.startswith(".")}


@pytest.mark.parametrize("nrows,B,S,id_codec,id_mode", [(0, 4, 2, "DENSE_INT", "sorted"), (1, 4, 2, "PFOR_INT", "sorted"), (9, 4, 2, "DENSE_INT", "random"),
                                                        (20, 4, 2, "PFOR_INT", "steps"), (5000, 64, 5, "PFOR_INT", "random"), (40_000, 64, 5, "DENSE_INT", "sorted"),
                                                        (9000, 1024, 3, "PFOR_INT", "steps"), (3 * (32 * 3 + 1), 32, 3, "PFOR_INT", "sorted")])
def test_oracle_writer_and_product_writer_agree_byte_for_byte(tmp_path, nrows, B, S, id_codec, id_mode):
    """Two independent restatements of SegmentWriter + LoaderCli + PFORCodecInt.encode (bit-stream accumulator in the
    product, word-indexed packing in the oracle) must write the same files."""
    a, b = tmp_path / "product", tmp_path / "oracle"
    name_arr = np.array([b"alice", b"bobby", b"carol"], dtype="S5")[np.arange(nrows) % 3]
    cols = make_table(a, "t", nrows, B, S, seed=3, id_codec=id_codec, id_mode=id_mode, extra_cols=[("name:DENSE_STRING:size=5", name_arr)])
    specs = [f"id:{id_codec}", "state:DENSE_STRING:size=2", "age:DENSE_TINYINT", "name:DENSE_STRING:size=5"]
    O.write_table(b, "t", specs, cols, B, S)
    fa, fb = _dir_bytes(a / "t"), _dir_bytes(b / "t")
    assert sorted(fa) == sorted(fb)
    for f in fa:
        assert fa[f] == fb[f], f


@pytest.mark.parametrize("codec", ["DENSE_INT", "PFOR_INT"])
def test_oracle_synth_writer_matches_product_synth_writer(tmp_path, codec):
    a, b = tmp_path / "product", tmp_path / "oracle"
    nrows, B, S = 3 * (64 * 7 + 1) + 100, 64, 7
    cid = L.CODEC_PFOR_INT if codec == "PFOR_INT" else L.CODEC_DENSE_INT
    synth_write(a, "s", nrows, B, S, cid, 0, -1, True)
    O.synth_write(b, "s", nrows, B, S, O.CODEC_PFOR_INT if codec == "PFOR_INT" else O.CODEC_DENSE_INT, 0, -1, True, nthreads=3)
    fa, fb = _dir_bytes(a / "s"), _dir_bytes(b / "s")
    assert sorted(fa) == sorted(fb) and len(fa) == 1 + 2 * 3 * 4
    for f in fa:
        assert fa[f] == fb[f], f
