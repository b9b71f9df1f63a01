"""CPU baseline = the reference's own way of doing ingestion, restated (TEST/BENCH
INFRASTRUCTURE ONLY): one Python call per SNP pair into a NumPy float32 array
(gretel/util.py:242-286 calling Hansel.add_observation), optionally with forked
workers over read chunks like util.py:294-326.  The matrix is banded instead of the
reference's dense (7,7,N+2,N+2) so that N=10k fits in host memory; the per-pair cost
(two dict lookups + one ndarray ``+= 1``) is the same.
"""
from __future__ import annotations

import multiprocessing as mp
import time

import numpy as np

SYMBOLS = ['A', 'C', 'G', 'T', 'N', '-', '_']


class BandHansel:
    def __init__(self, n_snps, W):
        self.N, self.W = n_snps, W
        self.m = np.zeros((n_snps + 2, W, 7, 7), dtype=np.float32)
        self.d = {s: i for i, s in enumerate(SYMBOLS)}

    def add_observation(self, a, b, i, j):
        self.m[j, j - i - 1, self.d[a], self.d[b]] += 1


def ingest(h, rank, off, codes, lo, hi):
    """util.py:226-286, literal, over reads [lo,hi)."""
    N = h.N
    slices = crumbs = covered = 0
    for r in range(lo, hi):
        seq = [SYMBOLS[c] for c in codes[off[r]:off[r + 1]]]
        if not len(seq) > 1:
            continue
        slices += 1
        rk = int(rank[r])
        support_len = len(seq)
        support_seq = "".join(seq)
        covered += len(support_seq.replace("N", "").replace("_", ""))
        for i in range(0, support_len):
            snp_a = support_seq[i]
            for j in range(i + 1, support_len):
                snp_b = support_seq[j]
                if snp_a in ['_', 'N']:
                    continue
                if i == 0 and j == 1 and rk == 0:
                    h.add_observation('_', snp_a, 0, 1)
                    h.add_observation(snp_a, snp_b, 1, 2)
                    crumbs += 1
                elif (j + rk + 1) == N and abs(i - j) == 1:
                    h.add_observation(snp_a, snp_b, N - 1, N)
                    h.add_observation(snp_b, '_', N, N + 1)
                    crumbs += 1
                else:
                    h.add_observation(snp_a, snp_b, i + rk + 1, j + rk + 1)
                    crumbs += 1
    return slices, crumbs, covered


_SHARED = {}


def _worker(args):
    lo, hi, N, W = args
    rank, off, codes = _SHARED["rank"], _SHARED["off"], _SHARED["codes"]   # inherited through fork
    h = BandHansel(N, W)
    t = time.perf_counter()
    res = ingest(h, rank, off, codes, lo, hi)
    return res, time.perf_counter() - t


def timed_ingest(rank, off, codes, n_snps, W, n_procs=1):
    """Ingest all given reads with ``n_procs`` forked workers (contiguous read chunks,
    private partial matrices).  Returns (crumbs, wall seconds)."""
    R = len(rank)
    t0 = time.perf_counter()
    _SHARED.update(rank=rank, off=off, codes=codes)
    if n_procs <= 1:
        (s, c, v), _ = _worker((0, R, n_snps, W))
        return c, time.perf_counter() - t0
    bounds = np.linspace(0, R, n_procs + 1).astype(int)
    ctx = mp.get_context("fork")
    with ctx.Pool(n_procs) as pool:
        out = pool.map(_worker, [(int(bounds[i]), int(bounds[i + 1]), n_snps, W) for i in range(n_procs)])
    crumbs = sum(o[0][1] for o in out)
    return crumbs, time.perf_counter() - t0
