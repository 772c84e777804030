"""Pins the CPU oracle: hand-derived known-answer vectors (SURVEY.md §8c — the reference ships no
tests or goldens and cannot be run without a JVM) plus round-trip / property checks of the
sorted-integer codec restatement (JavaFastPFOR is not in the reference tree: PARITY UNPINNED)."""
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle_lib as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "kat_vectors.json")
KAT = json.load(open(GOLDEN))


def test_dense_int_decode_is_little_endian():  # Conversions.scala:17-24 (Horner), DataType.scala:40-47
    for hexbytes, want in KAT["dense_int"]:
        assert O.lib().orc_bytes_to_int(bytes.fromhex(hexbytes)) == want
    buf = bytearray(4)
    import ctypes
    for v in (0, 1, -1, 67305985, -2**31, 2**31 - 1):
        b = ctypes.create_string_buffer(4)
        O.lib().orc_int_to_bytes(v, b)
        assert O.lib().orc_bytes_to_int(b.raw) == v
        assert b.raw == int(v & 0xFFFFFFFF).to_bytes(4, "little")


def test_dense_tinyint_and_string_decode():  # DataType.scala:61,70; DenseCodec.scala:48-74
    cells = O.dense_decode(bytes.fromhex("807f"), 1)
    assert list(np.frombuffer(cells, np.int8)) == [-128, 127]
    assert O.dense_decode(b"CANY", 2) == b"CANY"
    assert [O.dense_decode(b"CANY", 2)[i:i + 2] for i in (0, 2)] == [b"CA", b"NY"]


def test_dense_decode_ragged_tail_reuses_stale_chunk_bytes():  # DenseCodec.scala:41-44 read(chunk) semantics
    # 6 bytes at width 4: second read copies 2 bytes over the previous chunk
    out = O.dense_decode(bytes([1, 2, 3, 4, 9, 8]), 4)
    assert out == bytes([1, 2, 3, 4, 9, 8, 3, 4])


def test_jvm_narrowing():  # Select.scala:65,73,103,111,141,149
    for d, want in KAT["d2b"]:
        assert O.lib().orc_d2b(float(d)) == want
    for d, want in KAT["d2i"]:
        assert O.lib().orc_d2i(float(d)) == want
    assert O.lib().orc_d2i(float("nan")) == 0
    assert O.lib().orc_d2i(-7.9) == -7 and O.lib().orc_d2i(7.9) == 7
    assert O.lib().orc_d2i(-1e300) == -2**31


def test_pfor_known_answers_unverified_vs_real_javafastpfor():
    for case in KAT["pfor_blocks"]:
        vals = case["values"] if "values" in case else list(range(*case["range"]))
        assert O.pfor_encode(vals).hex() == case["hex"], case["name"]
        assert list(O.pfor_decode(bytes.fromhex(case["hex"]))) == vals


def test_pfor_sequential_1024_block_size():  # 1 + 8*(1+4) words + 8 pad bytes
    enc = O.pfor_encode(np.arange(1024))
    assert len(enc) == 4 * 41 + 8
    enc = O.pfor_encode(np.arange(1 << 20, (1 << 20) + 1024))
    # first mini-block needs 21 bits for the first delta (value itself vs base 0)
    assert len(enc) == 4 * (1 + 8 * 5 + 20) + 8


def test_pfor_unsorted_run_degrades_to_raw_32bit():
    vals = [5, 3] + list(range(10, 40))  # a negative delta sets bit 31 -> b = 32: raw copy
    w = O.iic_compress(vals)
    assert w[0] == 32 and w[1] == 32 and list(w[2:34]) == vals


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(-2**31, 2**31 - 1), min_size=0, max_size=700))
def test_pfor_roundtrip_arbitrary(vals):
    assert list(O.pfor_decode(O.pfor_encode(vals))) == vals


@settings(max_examples=100, deadline=None)
@given(st.integers(0, 2000), st.integers(0, 2**31 - 1), st.lists(st.integers(0, 70000), min_size=1, max_size=1300))
def test_pfor_roundtrip_sorted(n_unused, start, deltas):
    vals = (start + np.cumsum(deltas)).astype(np.int64).astype(np.int32)
    enc = O.pfor_encode(vals)
    assert len(enc) % 4 == 0 and enc[-8:] == b"\0" * 8
    assert np.array_equal(O.pfor_decode(enc), vals)
    # header: value count, big-endian
    assert int.from_bytes(enc[:4], "big") == len(vals)
