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
