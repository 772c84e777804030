"""Pinning the sorted-integer codec as far as this image allows (no JVM, no JavaFastPFOR jar):

* three independent restatements of PFORCodecInt.encode - the oracle (C, word-indexed packing), the product's host encoder
  (C++, bit-stream accumulator) and tests/pfor_independent.py (Python big integers, written from the upstream class
  structure) - must agree byte for byte on 10^4 random blocks, and every decoder must invert every encoder;
* hand-derived known answers on the independent implementation;
* real JavaFastPFOR vectors: `tools/gen_javafastpfor_goldens.scala` (run by anyone with a JVM and the 0.1.10 jar) writes
  tests/golden/javafastpfor_0.1.10.json; if that file is present all three encoders are checked against it and the
  codec's parity status changes from "unpinned" to pinned.
"""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
import pfor_independent as P
from immutable3_b200.loader import pfor_encode as product_encode

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "javafastpfor_0.1.10.json")


def _random_block(rng):
    n = int(rng.choice([0, 1, 2, 31, 32, 33, 63, 64, 127, 128, 129, 160, 255, 256, 300, 1000, 1024, 1025], p=[.02, .03, .03, .05, .08, .05, .05, .08, .05,
                                                                                                           .1, .05, .05, .05, .08, .08, .05, .08, .02]))
    kind = rng.integers(0, 7)
    if kind == 0:    # sorted, small steps
        v = np.cumsum(rng.integers(0, 4, size=n)) + int(rng.integers(-2**31, 2**31 - 5000))
    elif kind == 1:  # sorted, steps below 2^b
        v = np.cumsum(rng.integers(0, 1 << int(rng.integers(1, 31)), size=n, dtype=np.int64)) + int(rng.integers(-2**31, 2**31))
    elif kind == 2:  # constant
        v = np.full(n, int(rng.integers(-2**31, 2**31)), dtype=np.int64)
    elif kind == 3:  # unsorted
        v = rng.integers(-2**31, 2**31, size=n, dtype=np.int64)
    elif kind == 4:  # sequential ids (the benchmark's column)
        v = np.arange(n, dtype=np.int64) + int(rng.integers(0, 2**31 - 2000))
    elif kind == 5:  # mostly sorted with a few drops (negative deltas -> raw mini-blocks)
        v = np.cumsum(rng.integers(0, 100, size=n, dtype=np.int64))
        if n:
            v[rng.integers(0, n, size=max(1, n // 50))] -= 1_000_000
    else:            # extremes
        v = rng.choice(np.array([-2**31, 2**31 - 1, 0, -1, 1], dtype=np.int64), size=n)
    return ((v + 2**31) % 2**32 - 2**31).astype(np.int32)


def test_three_restatements_agree_on_ten_thousand_random_blocks():
    rng = np.random.default_rng(20261018)
    for i in range(10_000):
        v = _random_block(rng)
        a = O.pfor_encode(v)
        b = P.pfor_encode([int(x) for x in v])
        assert a == b, (i, len(v), v[:40])
        if i % 4 == 0:
            assert product_encode(v) == a, (i, len(v))
        if i % 8 == 0:
            assert P.pfor_decode(a) == [int(x) for x in v]
            assert np.array_equal(O.pfor_decode(b, cap=max(1, len(v)) + 8), v)


def test_known_answers_on_the_independent_restatement():
    assert P.iic_compress(list(range(32))) == [32, 1, 0xFFFFFFFE]
    assert P.pfor_encode(list(range(32))) == bytes.fromhex("00000020" "00000001" "FFFFFFFE") + bytes(8)
    assert P.iic_compress([300]) == [1, 0x0000822C]
    assert P.iic_compress([]) == [0]
    w = P.iic_compress(list(range(1000, 1128)))          # one super-block: first delta 1000 (10 bits), then ones
    assert w[0] == 128 and w[1] == (10 << 24) | (1 << 16) | (1 << 8) | 1 and len(w) == 2 + 10 + 3
    v = [5, 3] + [3] * 30                                # a negative delta sets bit 31: raw mini-block
    w = P.iic_compress(v)
    assert w[:2] == [32, 32] and w[2:] == v
    ws = [int(x) & 0xFFFFFFFF for x in O.iic_compress(np.array(list(range(0, 15 * 7, 7)), np.int32))]  # exploration/compression.sc:6 shape: 15 ints
    assert ws == P.iic_compress(list(range(0, 15 * 7, 7)))


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="no JavaFastPFOR vectors (tools/gen_javafastpfor_goldens.scala needs a JVM + the 0.1.10 jar)")
def test_real_javafastpfor_vectors():
    cases = json.load(open(GOLDEN))["cases"]
    assert len(cases) >= 10
    for c in cases:
        v = np.array(c["input"], dtype=np.int64).astype(np.int32)
        want = [int(x) & 0xFFFFFFFF for x in c["compressed"]]
        assert P.iic_compress([int(x) for x in v]) == want, c.get("name")
        assert [int(x) & 0xFFFFFFFF for x in O.iic_compress(v)] == want, c.get("name")
        assert product_encode(v) == b"".join(w.to_bytes(4, "big") for w in want) + bytes(8), c.get("name")


def test_golden_file_format_example_round_trips(tmp_path):
    """The generator's output format, exercised with vectors from our own encoder (NOT a pin - only keeps the loader honest)."""
    cases = [{"name": "seq32", "input": list(range(32)), "compressed": [32, 1, -2]}]
    p = tmp_path / "g.json"
    p.write_text(json.dumps({"library": "JavaFastPFOR 0.1.10", "class": "IntegratedIntCompressor", "cases": cases}))
    c = json.load(open(p))["cases"][0]
    assert [x & 0xFFFFFFFF for x in c["compressed"]] == P.iic_compress(c["input"])
