"""The C restatement (banded) must agree bit-for-bit with the literal Python oracle."""
import numpy as np
import pytest

from gretel_b200 import synth
from oracle import hansel_oracle as o


@pytest.mark.parametrize("seed", range(12))
def test_c_oracle_matches_python(c_oracle, seed):
    rng = np.random.default_rng(100 + seed)
    N = int(rng.integers(2, 16))
    R = int(rng.integers(1, 50))
    mk = int(rng.integers(2, 9))
    rank, off, codes = synth.random_packed(rng, N, R, mk, p_special=0.25, sort=bool(seed % 2))
    W = max(1, min(mk, N) - 1) if seed % 3 else N + 1
    v_site = "to" if seed % 4 == 0 else "from"
    h = o.OracleHansel.init_matrix(o.SYMBOLS, o.UNSYMBOLS, N, v_site=v_site)
    tot = o.ingest_packed(h, rank, off, codes, N)
    band, totals = c_oracle.ingest(rank, off, codes, N, W)
    assert tuple(totals[:3]) == tot
    assert int(band.sum()) == tot[1] + int(totals[3])
    assert np.array_equal(o.band_of(h, W), band.astype(np.float32))
    assert o.out_of_band_mass(h, W) == 0
    if tot[0] == 0:
        return
    h.L = int(rng.integers(1, 6))
    bf = band.astype(np.float32)
    bo = bf.copy()
    ho = h.copy()
    ca = c_oracle.counts_all(bf, N, W)
    for p in range(N + 1):
        d = h.get_counts_at(p)
        assert d["total"] == ca[p, 7]
        for i, s in enumerate(o.SYMBOLS):
            assert d.get(s, 0.0) == ca[p, i]
    for it in range(5):
        p, pr, mn = o.generate_path(N, h, ho)
        pc, res = c_oracle.generate_path(bf, bo, N, W, h.L, v_site=v_site)
        if p is None:
            assert pc is None
            break
        assert [o.CODE[s] for s in p] == list(pc)
        assert (pr["hp_current"], pr["hp_original"], mn) == res
        ratio = max(mn, 0.01)
        assert o.reweight_hansel_from_path(h, p, ratio) == c_oracle.reweight_path(bf, N, W, pc, ratio)
        assert np.array_equal(o.band_of(h, W), bf)


def test_c_oracle_rejects_bad_reads(c_oracle):
    rank = np.array([3], dtype=np.int32)
    off = np.array([0, 3], dtype=np.int64)
    codes = np.array([0, 1, 2], dtype=np.uint8)
    with pytest.raises(ValueError):
        c_oracle.ingest(rank, off, codes, 5, 4)       # rank + k > N
    with pytest.raises(ValueError):
        c_oracle.ingest(np.array([0], dtype=np.int32), off, codes, 5, 1)   # k-1 > W


def test_synthetic_workload_shape(c_oracle):
    w = synth.scaled(synth.WORKLOADS["hiv"], 4000)
    d = synth.generate(w)
    k = np.diff(d["off"])
    assert (k >= 2).all() and (np.diff(d["rank"]) >= 0).all()
    assert (d["rank"] + k <= w.n_snps).all()
    assert 20 < k.mean() < 32                      # ~26 SNPs per 250 bp read at 1000 SNPs / 9.7 kb
    band, totals = c_oracle.ingest(d["rank"], d["off"], d["codes"], w.n_snps, d["max_k"] - 1)
    assert totals[0] == len(k)
    assert totals[1] <= d["n_pairs"]               # pairs whose first allele is N carry no crumb
    assert int(band.sum()) == totals[1] + totals[3]
    # same seed => same reads; another shard => different reads of the same strains
    d2 = synth.generate(w)
    assert np.array_equal(d2["codes"], d["codes"])
    d3 = synth.generate(w, shard=1)
    assert np.array_equal(d3["strains"], d["strains"]) and not np.array_equal(d3["rank"], d["rank"])
