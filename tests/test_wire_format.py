"""Host-side encoders of the wire formats (no GPU): decoding them with numpy gives the packed reads back."""
import os

import numpy as np
import pytest

from gretel_b200 import synth, util


def _decode_dense(d):
    delta = d.rank_delta.astype(np.int64)
    delta[d.esc_idx] = d.esc_delta
    rank = np.cumsum(delta).astype(np.int32)
    off = np.zeros(d.n_reads + 1, np.int64)
    np.cumsum(d.klen.astype(np.int64), out=off[1:])
    f = np.stack([(d.codes2 >> s) & 3 for s in (0, 2, 4, 6)], axis=1).reshape(-1)[:d.n_codes].astype(np.uint8)
    f[d.exc_pos] += 4
    return rank, off, f


@pytest.mark.parametrize("seed", range(5))
def test_dense_packed_round_trip(seed):
    rng = np.random.default_rng(seed)
    N = [8, 50, 3000, 20000, 400][seed]
    rank, off, codes = synth.random_packed(rng, N, int(rng.integers(1, 300)), [4, 11, 9, 7, 300][seed],
                                           p_special=0.2)
    d = util.dense_packed(rank, off, codes)
    r2, o2, c2 = _decode_dense(d)
    assert np.array_equal(r2, rank) and np.array_equal(o2, off) and np.array_equal(c2, codes)
    assert d.klen.dtype == (np.uint16 if np.diff(off).max() > 255 else np.uint8)
    assert d.nbytes < rank.nbytes + off.nbytes + codes.nbytes
    pieces = util.dense_chunks(rank, off, codes, 3)
    assert sum(p.n_reads for p in pieces) == len(rank) and sum(p.n_codes for p in pieces) == len(codes)
    got = [_decode_dense(p) for p in pieces]
    assert np.array_equal(np.concatenate([g[0] for g in got]), rank)
    assert np.array_equal(np.concatenate([g[2] for g in got]), codes)


def test_dense_packed_needs_sorted_reads():
    with pytest.raises(ValueError):
        util.dense_packed(np.array([2, 1], np.int32), np.array([0, 2, 4], np.int64), np.zeros(4, np.uint8))


def test_compact_packed_round_trip():
    rng = np.random.default_rng(9)
    rank, off, codes = synth.random_packed(rng, 40, 101, 9, p_special=0.2)
    klen, codes4, n = util.compact_packed(off, codes)
    assert n == len(codes) and np.array_equal(np.cumsum(klen), off[1:])
    back = np.stack([codes4 & 15, codes4 >> 4], axis=1).reshape(-1)[:n]
    assert np.array_equal(back, codes)


@pytest.mark.parametrize("seed", range(5))
def test_native_dense_encoder_matches_numpy(seed):
    """hx_dense_encode (C++, threaded) produces the same bytes as util.dense_packed."""
    rng = np.random.default_rng(100 + seed)
    N = [8, 50, 3000, 20000, 400][seed]
    n_reads = [7, 300, 150, 200_000, 90][seed]          # 200k reads: several encoder threads
    rank, off, codes = synth.random_packed(rng, N, n_reads, [4, 11, 9, 7, 300][seed], p_special=0.2)
    a = util.dense_packed(rank, off, codes)
    b = util.dense_packed_native(rank, off, codes, n_threads=4)
    assert (a.n_reads, a.n_codes) == (b.n_reads, b.n_codes)
    for x, y in zip(a.arrays(), b.arrays()):
        assert x.dtype == y.dtype and np.array_equal(x, y)
    assert np.array_equal(a.blob, b.blob)
    # a chunk that starts in the middle of the allele stream (off[0] != 0)
    if len(rank) > 4:
        h = len(rank) // 2
        a2 = util.dense_packed(rank[h:], off[h:], codes)
        b2 = util.dense_packed_native(rank[h:], off[h:], codes, n_threads=3)
        assert np.array_equal(a2.blob, b2.blob)


def test_native_dense_encoder_rejects_bad_input():
    with pytest.raises(ValueError):
        util.dense_packed_native(np.array([2, 1], np.int32), np.array([0, 2, 4], np.int64), np.zeros(4, np.uint8))
    with pytest.raises(ValueError):
        util.dense_packed_native(np.array([0, 1], np.int32), np.array([0, 2, 4], np.int64), np.array([0, 9, 1, 2], np.uint8))


def test_dense_encoders_agree_on_random_shapes():
    """Property check over many small random inputs (empty read lists, reads of 0 or 1 SNPs, all-rare reads,
    huge rank gaps, allele streams whose length is not a multiple of 4 or 16): numpy and C++ encoders produce the
    same bytes and decode back to the input."""
    rng = np.random.default_rng(2026)
    for trial in range(150):
        n_reads = int(rng.integers(0, 40))
        N = int(rng.choice([3, 17, 300, 70000]))
        max_k = int(rng.integers(0, min(N, 9) + 1))
        rank, off, codes = synth.random_packed(rng, N, n_reads, max_k, p_special=float(rng.choice([0.0, 0.3, 1.0])))
        a = util.dense_packed(rank, off, codes)
        b = util.dense_packed_native(rank, off, codes, n_threads=int(rng.integers(1, 4)))
        assert np.array_equal(a.blob, b.blob), trial
        r2, o2, c2 = _decode_dense(a)
        assert np.array_equal(r2, rank) and np.array_equal(o2, off) and np.array_equal(c2, codes), trial


def _simd_levels_agree():
    rng = np.random.default_rng(77)
    rank, off, codes = synth.random_packed(rng, 5000, 150_000, 30, p_special=0.02)
    codes[rng.integers(0, len(codes), 50)] = 6
    blobs = []
    for level in ("0", "1", "2"):          # 64-bit words, SSE4.1, AVX2 (a level the CPU lacks falls back to the one below)
        os.environ["HX_DENSE_SIMD"] = level
        try:
            blobs.append(util.dense_packed_native(rank, off, codes, n_threads=3).blob.copy())
            h = len(rank) // 3                # a chunk starting at an allele offset that is not a multiple of 32
            blobs.append(util.dense_packed_native(rank[h:], off[h:], codes, n_threads=2).blob.copy())
            bad = codes.copy()
            bad[len(bad) // 2 + 5] = 9         # an invalid allele must be caught at every level
            with pytest.raises(ValueError):
                util.dense_packed_native(rank, off, bad, n_threads=3)
        finally:
            del os.environ["HX_DENSE_SIMD"]
    assert np.array_equal(blobs[0], util.dense_packed(rank, off, codes).blob)
    for i in (2, 4):
        assert np.array_equal(blobs[i], blobs[0]) and np.array_equal(blobs[i + 1], blobs[1])


def test_dense_encoder_simd_levels_agree():
    """The SSE4.1 / AVX2 allele packers write the same bytes as the 64-bit loop and the numpy encoder."""
    _simd_levels_agree()


@pytest.mark.gpu
def test_dense_encoder_simd_levels_agree_on_the_gpu_box():
    """The same on the GPU box's host CPU (it may have AVX2 where the build container does not)."""
    _simd_levels_agree()
