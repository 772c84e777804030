"""A second, independent restatement of the sorted-integer codec, used only to cross-check the oracle.

PFORCodecInt.encode (PFORCodec.scala:17-28) = `new IntegratedIntCompressor().compress(ints)` of JavaFastPFOR 0.1.10
(project/Dependencies.scala:4), each word written with ByteBuffer.putInt (big-endian), the whole (4*words + 8)-byte
backing array emitted.  The library is not under /root/reference; this file restates its PUBLISHED algorithm from the
structure of the upstream Java classes - written without reference to oracle/oracle.c or csrc/writer.cpp, and in a
different formulation (Python big integers as bit streams) so that a shared misunderstanding would have to be shared
three times:

    IntegratedIntCompressor.compress        out[0] = n; codec.headlessCompress(in, 0, n, out, 1, initvalue = 0)
    SkippableIntegratedComposition          F1 = IntegratedBinaryPacking on the first n - n % 32 ints, F2 =
                                            IntegratedVariableByte on the rest, sharing one running `initvalue`
    IntegratedBinaryPacking.headlessCompress
        while >= 128 ints remain: bits b1..b4 = Util.maxdiffbits of four 32-int runs (each against the last value of the
        previous run), one header word (b1<<24 | b2<<16 | b3<<8 | b4), then the four packed runs (b_i words each);
        while >= 32 remain: one header word b, then b words
    IntegratedBitPacking.integratedpack     32 deltas of `bit` bits, little-endian bit order; bit == 32 copies the VALUES
    IntegratedVariableByte.headlessCompress deltas as 7-bit groups, least significant first, the LAST byte of a value
                                            carries bit 7; bytes in a LITTLE_ENDIAN buffer, zero-padded to 4, read as ints

Byte compatibility with the real jar remains unverified offline (tools/gen_javafastpfor_goldens.scala produces the vectors
that would pin it; tests/test_codec_pinning.py loads them if present).
"""
from __future__ import annotations

from typing import List, Sequence

M32 = 0xFFFFFFFF


def _u32(x: int) -> int:
    return x & M32


def _s32(x: int) -> int:
    x &= M32
    return x - (1 << 32) if x & 0x80000000 else x


def maxdiffbits(initoffset: int, run: Sequence[int]) -> int:
    """Util.maxdiffbits: bit length of the OR of the (wrapping) deltas of a 32-int run."""
    mask, prev = 0, initoffset
    for v in run:
        mask |= _u32(v - prev)
        prev = v
    return mask.bit_length()


def integratedpack(initoffset: int, run: Sequence[int], bit: int) -> List[int]:
    if bit == 0:
        return []
    if bit == 32:
        return [_u32(v) for v in run]
    stream, prev = 0, initoffset
    for j, v in enumerate(run):
        stream |= _u32(v - prev) << (j * bit)   # deltas fit `bit` bits by construction
        prev = v
    return [(stream >> (32 * w)) & M32 for w in range(bit)]


def integratedunpack(initoffset: int, words: Sequence[int], bit: int) -> List[int]:
    if bit == 0:
        return [_s32(initoffset)] * 32
    if bit == 32:
        return [_s32(w) for w in words[:32]]
    stream = 0
    for w, word in enumerate(words[:bit]):
        stream |= (word & M32) << (32 * w)
    out, prev = [], initoffset
    for j in range(32):
        prev = _u32(prev + ((stream >> (j * bit)) & ((1 << bit) - 1)))
        out.append(_s32(prev))
    return out


def variable_byte(initoffset: int, values: Sequence[int]) -> List[int]:
    raw = bytearray()
    prev = initoffset
    for v in values:
        val = _u32(v - prev)
        prev = v
        while val >= 0x80:
            raw.append(val & 0x7F)
            val >>= 7
        raw.append(val | 0x80)
    while len(raw) % 4:
        raw.append(0)
    return [int.from_bytes(raw[i:i + 4], "little") for i in range(0, len(raw), 4)]


def iic_compress(values: Sequence[int]) -> List[int]:
    """IntegratedIntCompressor.compress -> list of unsigned 32-bit words."""
    vals = [_s32(v) for v in values]
    n = len(vals)
    out = [_u32(n)]
    if n == 0:
        return out
    init = 0
    packed = n - n % 32
    s = 0
    while s + 128 <= packed:
        inits = [init, vals[s + 31], vals[s + 63], vals[s + 95]]
        bits = [maxdiffbits(inits[k], vals[s + 32 * k: s + 32 * k + 32]) for k in range(4)]
        out.append((bits[0] << 24) | (bits[1] << 16) | (bits[2] << 8) | bits[3])
        for k in range(4):
            out += integratedpack(inits[k], vals[s + 32 * k: s + 32 * k + 32], bits[k])
        init = vals[s + 127]
        s += 128
    while s < packed:
        b = maxdiffbits(init, vals[s:s + 32])
        out.append(b)
        out += integratedpack(init, vals[s:s + 32], b)
        init = vals[s + 31]
        s += 32
    if n > packed:
        out += variable_byte(init, vals[packed:])
    return out


def iic_uncompress(words: Sequence[int]) -> List[int]:
    n = words[0] & M32
    out: List[int] = []
    pos, init = 1, 0
    packed = n - n % 32
    while len(out) + 128 <= packed:
        h = words[pos] & M32
        pos += 1
        for k in range(4):
            b = (h >> (24 - 8 * k)) & 0xFF
            run = integratedunpack(init, words[pos:pos + b], b)
            pos += b
            out += run
            init = _u32(run[-1])
    while len(out) < packed:
        b = words[pos] & M32
        pos += 1
        run = integratedunpack(init, words[pos:pos + b], b)
        pos += b
        out += run
        init = _u32(run[-1])
    raw = b"".join(int(w & M32).to_bytes(4, "little") for w in words[pos:])
    i = 0
    while len(out) < n:
        val, shift = 0, 0
        while True:
            c = raw[i]
            i += 1
            val |= (c & 0x7F) << shift
            shift += 7
            if c & 0x80:
                break
        init = _u32(init + val)
        out.append(_s32(init))
    return out


def pfor_encode(values: Sequence[int]) -> bytes:
    """PFORCodecInt.encode: ByteBuffer.allocate(words * 4 + 8), putInt per word (big-endian), whole array."""
    words = iic_compress(values)
    return b"".join(w.to_bytes(4, "big") for w in words) + bytes(8)


def pfor_decode(data: bytes) -> List[int]:
    body = data[:-8]
    return iic_uncompress([int.from_bytes(body[i:i + 4], "big") for i in range(0, len(body), 4)])
