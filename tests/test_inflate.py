"""The packer's raw-DEFLATE decoder (csrc/hx_inflate.h) against zlib: same bytes out for everything zlib can write,
a clean refusal for everything zlib refuses."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

from gretel_b200 import _lib


def _inflate(payload, m, slack=8, use_zlib=0):
    lib = _lib.load()
    src = np.frombuffer(payload + b"\xa5" * slack, np.uint8).copy()
    dst = np.full(m + 64, 0xEE, np.uint8)                  # canary behind the output
    rc = lib.hx_inflate_raw(src.ctypes.data, len(payload), len(payload) + slack, dst.ctypes.data, m, use_zlib)
    assert np.all(dst[m:] == 0xEE), "wrote past the end of the output"
    return rc, dst[:m].tobytes()


def _deflate(data, level, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=-15, memlevel=8):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, memlevel, strategy)
    return c.compress(data) + c.flush()


def _samples():
    rng = np.random.default_rng(7)
    yield b""
    yield b"a"
    yield b"ab" * 40000                                    # period-2 matches
    yield b"\x00" * 65536                                  # distance-1 runs
    yield bytes(rng.integers(0, 256, 65536, dtype=np.uint8))        # incompressible -> stored blocks
    yield bytes(rng.integers(0, 4, 65280, dtype=np.uint8))          # 2-bit alphabet
    yield bytes(rng.integers(33, 74, 65280, dtype=np.uint8))        # base-quality-like literals
    text = b" ".join(b"read%07d\t%d\tctg\t%d\t60\t150M\t*\t0\t0" % (i, 99 if i & 1 else 147, 1000 + 3 * i) for i in range(1500))
    yield text[:65280]
    # BAM-like records: fixed header fields, names, 4-bit bases, qualities
    rec = bytearray()
    for i in range(400):
        rec += (200).to_bytes(4, "little") + (0).to_bytes(4, "little") + (1000 + 7 * i).to_bytes(4, "little")
        rec += b"read%09d\0" % i + bytes(rng.integers(0, 256, 75, dtype=np.uint8)) + bytes(rng.integers(30, 42, 150, dtype=np.uint8))
    yield bytes(rec[:65280])
    for n in (1, 2, 3, 7, 8, 9, 257, 258, 259, 300, 1000, 4095):   # short outputs: the careful tail path only
        yield bytes(rng.integers(0, 3, n, dtype=np.uint8))
    # skewed alphabets give code lengths up to 15 (second-level tables)
    p = 0.5 ** np.arange(1, 41); p = np.concatenate([p, np.full(216, (1 - p.sum()) / 216)])
    yield bytes(rng.choice(256, size=65000, p=p / p.sum()).astype(np.uint8))


@pytest.mark.parametrize("level", [0, 1, 4, 6, 9])
def test_matches_zlib_on_everything_zlib_writes(level):
    n = 0
    for data in _samples():
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
            payload = _deflate(data, level, strategy)
            for slack in (8, 0, 3):
                rc, out = _inflate(payload, len(data), slack)
                assert rc == 0 and out == data, (level, strategy, len(data), slack)
            n += 1
    assert n > 50


def test_multi_block_and_flush_points():
    rng = np.random.default_rng(3)
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    parts, data = [], b""
    for i in range(12):
        piece = bytes(rng.integers(0, 1 + 3 * (i % 5) ** 3, 3000 + 997 * i, dtype=np.uint8))
        data += piece
        parts.append(c.compress(piece) + c.flush(zlib.Z_FULL_FLUSH if i % 3 else zlib.Z_SYNC_FLUSH))   # empty stored blocks in between
    payload = b"".join(parts) + c.flush()
    rc, out = _inflate(payload, len(data))
    assert rc == 0 and out == data


def test_refuses_what_zlib_refuses():
    rng = np.random.default_rng(11)
    data = bytes(rng.integers(0, 6, 20000, dtype=np.uint8)) + b"tail" * 100
    payload = _deflate(data, 6)
    assert _inflate(payload, len(data))[0] == 0
    assert _inflate(payload, len(data) - 1)[0] != 0          # inflates to more than announced
    assert _inflate(payload, len(data) + 1)[0] != 0          # ... to less
    assert _inflate(payload[:-3], len(data))[0] != 0         # truncated stream (reads into the slack, must notice)
    assert _inflate(payload[: len(payload) // 2], len(data))[0] != 0
    bad = 0
    for trial in range(300):                                  # random corruption: never a crash, never a wrong "ok"
        b = bytearray(payload)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        rc, out = _inflate(bytes(b), len(data))
        rz, outz = _inflate(bytes(b), len(data), use_zlib=1)
        if rc == 0:
            assert rz == 0 and out == outz, trial             # accepted -> zlib accepts it too, same bytes
        else:
            bad += 1
    assert bad > 100
    for trial in range(200):                                  # pure noise
        noise = bytes(rng.integers(0, 256, int(rng.integers(1, 400)), dtype=np.uint8))
        rc, out = _inflate(noise, 1000)
        if rc == 0:
            assert _inflate(noise, 1000, use_zlib=1) == (0, out)


def test_reference_fixture_blocks():
    """Every BGZF block of the reference's own test BAM (tests/golden/ref_test.bam)."""
    path = os.path.join(os.path.dirname(__file__), "golden", "ref_test.bam")
    raw = open(path, "rb").read()
    p, n = 0, 0
    while p < len(raw):
        xlen = int.from_bytes(raw[p + 10:p + 12], "little")
        bsize = int.from_bytes(raw[p + 16:p + 18], "little") + 1
        payload = raw[p + 12 + xlen:p + bsize - 8]
        isize = int.from_bytes(raw[p + bsize - 4:p + bsize], "little")
        rc, out = _inflate(payload, isize)
        assert rc == 0 and out == zlib.decompress(payload, -15) and zlib.crc32(out) == int.from_bytes(raw[p + bsize - 8:p + bsize - 4], "little")
        p += bsize
        n += 1
    assert n >= 2


class _Bits:
    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def put(self, v, nbits):                   # LSB first
        self.acc |= v << self.n
        self.n += nbits
        while self.n >= 8:
            self.out.append(self.acc & 255)
            self.acc >>= 8
            self.n -= 8

    def code(self, c, nbits):                  # a Huffman code: most significant bit first
        self.put(int(format(c, "0%db" % nbits)[::-1], 2), nbits)

    def done(self):
        if self.n:
            self.out.append(self.acc & 255)
        return bytes(self.out)


def _canonical(lens):
    codes, code = {}, 0
    for ln in range(1, 16):
        for s in sorted(k for k, v in lens.items() if v == ln):
            codes[s] = (code, ln)
            code += 1
        code <<= 1
    return codes


_LBASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
_LEXTRA = [0] * 8 + [1] * 4 + [2] * 4 + [3] * 4 + [4] * 4 + [5] * 4 + [0]
_DBASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
          8193, 12289, 16385, 24577]
_DEXTRA = [0, 0, 0, 0] + [i // 2 for i in range(2, 28)]


def test_hand_built_block_with_15_bit_codes():
    """A dynamic block written by hand whose literal/length AND distance codes are maximally skewed (lengths 1, 2, ...,
    14, 15, 15): every code longer than the primary tables (11 / 8 bits) goes through a second-level table.  zlib must
    accept the stream (it is a valid, complete code) and both decoders must agree byte for byte."""
    rng = np.random.default_rng(5)
    lit_syms = [ord(c) for c in "etaoinshrdlu"] + [256, 257, 264, 285]            # 12 literals, EOB, three length codes
    for trial in range(6):
        order = list(rng.permutation(16))
        ll = {lit_syms[order[i]]: min(i + 1, 15) for i in range(16)}               # lengths 1..15, 15: a complete code
        dorder = list(rng.permutation(16))
        dl = {int(dorder[i]) * 29 // 15: min(i + 1, 15) for i in range(16)}         # 16 distinct distance codes in 0..29
        assert len(dl) == 16
        lc, dc = _canonical(ll), _canonical(dl)
        w = _Bits()
        w.put(1, 1); w.put(2, 2)                                                    # BFINAL, dynamic
        hlit, hdist = max(ll) + 1, max(dl) + 1
        w.put(hlit - 257, 5); w.put(hdist - 1, 5); w.put(19 - 4, 4)
        cl_order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
        for s in cl_order:                                                          # code-length code: 0..15 at 4 bits each
            w.put(4 if s < 16 else 0, 3)
        for s in range(hlit):
            w.code(ll.get(s, 0), 4)
        for s in range(hdist):
            w.code(dl.get(s, 0), 4)
        # symbols: literals with every code length, matches with every length / distance code that is legal here
        out = bytearray()
        lits = [s for s in ll if s < 256]
        for step in range(4000):
            if step % 3 == 2 and len(out) > 0:
                lsym = int(rng.choice([257, 264, 285]))
                dsym = int(rng.choice([d for d in dl if _DBASE[d] <= len(out)] or [-1]))
                if dsym >= 0:
                    length = _LBASE[lsym - 257] + int(rng.integers(0, 1 << _LEXTRA[lsym - 257]))
                    dist = min(len(out), _DBASE[dsym] + int(rng.integers(0, 1 << _DEXTRA[dsym])))
                    dist = max(dist, _DBASE[dsym])
                    w.code(*lc[lsym]); w.put(length - _LBASE[lsym - 257], _LEXTRA[lsym - 257])
                    w.code(*dc[dsym]); w.put(dist - _DBASE[dsym], _DEXTRA[dsym])
                    for _ in range(length):
                        out.append(out[-dist])
                    continue
            s = int(rng.choice(lits))
            w.code(*lc[s])
            out.append(s)
        w.code(*lc[256])
        stream = w.done()
        assert zlib.decompress(stream, -15) == bytes(out)                           # the encoder above is right
        for slack in (8, 0):
            rc, got = _inflate(stream, len(out), slack)
            assert rc == 0 and got == bytes(out), (trial, slack)
