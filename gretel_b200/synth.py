"""Synthetic packed-read generators for BASELINE.json configs 2-5.

All draws come from ``numpy.random.Generator(PCG64(seed))`` with
``seed = 20260000 + config#`` (+ rank for per-GPU shards), so the oracle, the tests
and the bench see identical inputs.  Output is the packed format the CUDA path
consumes directly: ``rank int32[R]`` (0-based index of the first SNP on the read),
``off int64[R+1]``, ``codes uint8[sum k]`` with codes A0 C1 G2 T3 N4 -5 _6
(order fixed by gretel/util.py:83); reads are sorted by start (=> by rank) and reads
covering fewer than two SNPs are dropped (util.py:230).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

ABUNDANCES = np.array([0.40, 0.25, 0.15, 0.12, 0.08])


@dataclass(frozen=True)
class Workload:
    name: str
    config: int          # BASELINE.json configs[] index
    genome_len: int      # G
    n_snps: int          # N
    n_reads: int         # R (before dropping reads with <2 SNPs)
    read_len: int        # mean length for long reads
    eps: float           # substitution rate per covered SNP
    delta: float         # deletion rate per covered SNP
    long_reads: bool = False


WORKLOADS = {
    # C2: synthetic HIV-1-like 9.7 kb contig, 5 strains, ~1k SNPs, 200k x 250 bp reads
    "hiv": Workload("hiv", 1, 9_700, 1_000, 200_000, 250, 0.005, 0.001),
    # C3: synthetic metagenomic gene region, 10k SNPs, 10M x 150 bp reads
    "metagenome": Workload("metagenome", 2, 100_000, 10_000, 10_000_000, 150, 0.005, 0.001),
    # C4: 100k x 10 kb ONT-like reads, ~300 SNPs/read
    "ont": Workload("ont", 3, 333_000, 10_000, 100_000, 10_000, 0.05, 0.02, long_reads=True),
}


def scaled(w: Workload, n_reads: int) -> Workload:
    return Workload(w.name, w.config, w.genome_len, w.n_snps, int(n_reads), w.read_len,
                    w.eps, w.delta, w.long_reads)


def make_strains(rng, n_snps, n_strains=5):
    """Per site 2-3 distinct alleles from ACGT, assigned so that >=2 strains differ."""
    n_alleles = rng.integers(2, 4, size=n_snps)                   # 2 or 3
    perm = np.argsort(rng.random((n_snps, 4)), axis=1)            # random allele order per site
    assign = rng.integers(0, 3, size=(n_strains, n_snps)) % n_alleles[None, :]
    # force strains 0 and 1 onto different alleles so every site is polymorphic
    assign[0, :] = 0
    assign[1, :] = 1
    strains = np.take_along_axis(perm.T, assign, axis=0)          # [S][N] codes 0..3
    return np.ascontiguousarray(strains, dtype=np.uint8)


def make_sites(rng, genome_len, n_snps):
    return np.sort(rng.choice(genome_len, size=n_snps, replace=False) + 1).astype(np.int64)


def generate(w: Workload, seed=None, shard=0, chunk=2_000_000):
    """Return dict(rank, off, codes, n_snps, sites, strains, max_k, n_obs_upper)."""
    if seed is None:
        seed = 20260000 + w.config + 1
    # strains/sites depend on the workload seed only; reads also on the shard so that
    # every GPU of a weak-scaling run ingests a different slice of the same metagenome.
    rng_g = np.random.Generator(np.random.PCG64(seed))
    sites = make_sites(rng_g, w.genome_len, w.n_snps)
    strains = make_strains(rng_g, w.n_snps)
    rng = np.random.Generator(np.random.PCG64([seed, 7919 + shard]))
    R, G = w.n_reads, w.genome_len
    if w.long_reads:
        sigma = 0.3
        mu = np.log(w.read_len) - 0.5 * sigma * sigma
        lens = np.clip(rng.lognormal(mu, sigma, size=R), 2_000, 30_000).astype(np.int64)
    else:
        lens = np.full(R, w.read_len, dtype=np.int64)
    start = rng.integers(1 - w.read_len + 1, G + 1, size=R)
    order = np.argsort(start, kind="stable")
    start = start[order]
    lens = lens[order]
    s = np.maximum(start, 1)
    e = np.minimum(start + lens - 1, G)
    rank = np.searchsorted(sites, s, side="left")
    hi = np.searchsorted(sites, e, side="right")
    k = hi - rank
    keep = k >= 2
    rank, k = rank[keep].astype(np.int64), k[keep].astype(np.int64)
    strain_of = rng.choice(len(ABUNDANCES), size=R, p=ABUNDANCES)[keep].astype(np.int64)
    off = np.zeros(len(k) + 1, dtype=np.int64)
    np.cumsum(k, out=off[1:])
    total = int(off[-1])
    codes = np.empty(total, dtype=np.uint8)
    flat_strains = strains.reshape(-1)
    N = w.n_snps
    # chunk the flat expansion to bound temporaries (150 M codes for C3)
    r0 = 0
    nreads = len(k)
    while r0 < nreads:
        r1 = min(nreads, r0 + chunk)
        kk = k[r0:r1]
        n = int(off[r1] - off[r0])
        rep = np.repeat(np.arange(r1 - r0), kk)
        t = np.arange(n, dtype=np.int64) - np.repeat(off[r0:r1] - off[r0], kk)
        site = rank[r0:r1][rep] + t
        c = flat_strains[strain_of[r0:r1][rep] * N + site]
        u = rng.random(n, dtype=np.float32)
        sub = u < w.eps
        if sub.any():
            shift = rng.integers(1, 4, size=int(sub.sum())).astype(np.uint8)
            c[sub] = (c[sub] + shift) % 4
        c[(u >= w.eps) & (u < w.eps + w.delta)] = 5                # '-'
        c[(u >= w.eps + w.delta) & (u < w.eps + w.delta + 0.001)] = 4   # 'N'
        codes[off[r0]:off[r1]] = c
        r0 = r1
    return {
        "rank": rank.astype(np.int32), "off": off, "codes": codes, "n_snps": N,
        "sites": sites, "strains": strains, "max_k": int(k.max()) if len(k) else 0,
        "n_pairs": int((k * (k - 1) // 2).sum()),
    }


def random_packed(rng, n_snps, n_reads, max_k, p_special=0.1, sort=True):
    """Small adversarial packed reads for parity tests: any code 0..6, k in [0, max_k]."""
    ks = rng.integers(0, max_k + 1, size=n_reads)
    ks = np.minimum(ks, n_snps)
    ranks = np.array([rng.integers(0, n_snps - k + 1) for k in ks], dtype=np.int64)
    if sort:
        o = np.argsort(ranks, kind="stable")
        ranks, ks = ranks[o], ks[o]
    off = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(ks, out=off[1:])
    codes = rng.integers(0, 4, size=int(off[-1])).astype(np.uint8)
    special = rng.random(len(codes)) < p_special
    codes[special] = rng.integers(4, 7, size=int(special.sum())).astype(np.uint8)
    return ranks.astype(np.int32), off, codes
