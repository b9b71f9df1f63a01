"""Pins the oracle to the reference's own golden vectors (tests/test_test.py under
/root/reference).  ref_test.bam / ref_test.vcf.gz are byte copies of the reference's
binary test fixtures tests/data/test.bam and tests/data/test.vcf.gz."""
import os

import numpy as np

from gretel_b200 import bamio
from oracle import hansel_oracle as o

EXP = [1, 2, 10]


def _load(golden_dir, end):
    vcf = bamio.process_vcf(os.path.join(golden_dir, "ref_test.vcf.gz"), "hoot", 1, end)
    return vcf


def test_vcf(golden_dir):
    # tests/test_test.py:19-31
    v = _load(golden_dir, 19)
    assert v["N"] == 3
    assert v["snp_rev"] == {0: 1, 1: 2, 2: 10}
    assert v["snp_fwd"] == {1: 0, 2: 1, 10: 2}
    assert len(v["region"]) == 19 + 1
    for i in range(len(v["region"])):
        assert v["region"][i] == (1 if i in EXP else 0)


def test_bam_known_answers(golden_dir):
    # tests/test_test.py:33-52 (thread count does not exist on this path: the union of the
    # reference's windows is what the packer emits)
    v = _load(golden_dir, 20)
    rank, off, codes = bamio.pack_bam(os.path.join(golden_dir, "ref_test.bam"), "hoot", 1, 20, v)
    h = o.load_from_packed(rank, off, codes, v["N"])
    assert h.n_slices == 5
    assert h.n_crumbs == 9
    assert h.L > 0
    g = h.get_observation
    assert g('_', 'A', 0, 1) == 1
    assert g('A', 'A', 1, 2) == 1
    assert g('A', 'A', 1, 3) == 1
    assert g('A', 'A', 1, 4) == 0
    assert g('C', 'C', 1, 2) == 1
    assert g('C', 'C', 1, 3) == 1
    assert g('C', 'C', 1, 4) == 0
    assert g('T', 'T', 1, 2) == 2
    assert g('G', 'G', 1, 2) == 0
    assert g('G', 'G', 2, 3) == 0
    assert g('G', 'G', 3, 4) == 1
    assert g('G', '_', 4, 5) == 1
    # SURVEY.md section 4 hand trace: 9 crumbs + 4 start sentinels + 1 end sentinel, L = ceil(12/5)
    assert h.m.sum() == 14
    assert h.L == 3


def test_packer_matches_sam_text(golden_dir):
    """The BAM decoder agrees with the human-readable SAM the reference ships beside it."""
    refs, recs = bamio.read_bam(os.path.join(golden_dir, "ref_test.bam"))
    recs = list(recs)
    sam = [l.rstrip("\n").split("\t") for l in open(os.path.join(golden_dir, "ref_test.sam")) if l[0] != "@"]
    assert refs == [("hoot", 20), ("meow", 20)]
    assert len(recs) == len(sam) == 6
    for r, s in zip(recs, sam):
        assert r.name == s[0] and r.flag == int(s[1]) and refs[r.tid][0] == s[2]
        assert r.pos + 1 == int(s[3]) and r.seq == s[9]
        assert "".join("%d%s" % (l, "MIDNSHP=X"[op]) for op, l in r.cigar) == s[5]


def test_counts_and_gap_check(golden_dir):
    v = _load(golden_dir, 20)
    rank, off, codes = bamio.pack_bam(os.path.join(golden_dir, "ref_test.bam"), "hoot", 1, 20, v)
    h = o.load_from_packed(rank, off, codes, v["N"])
    assert h.get_counts_at(0) == {'_': 4.0, 'total': 4.0}
    assert h.get_counts_at(1) == {'A': 1.0, 'C': 1.0, 'T': 2.0, 'total': 4.0}
    assert o.gap_check(h, v["N"]) == []


def test_reweight_closed_form():
    """gretel.py:79-98 visits: pairs (p,q), q<=N-1 once; adjacent (i,i+1), i<=N-2 twice;
    (N-1,N) once; (p,N), p<=N-2 never; (N,N+1) once with ('path[N]','_')."""
    for N in (1, 2, 3, 4, 6, 9):
        class Rec:
            def __init__(self): self.calls = []
            def reweight_observation(self, a, b, i, j, r):
                self.calls.append((a, b, i, j)); return 0.0
        rec = Rec()
        path = ['_'] + ["s%d" % i for i in range(1, N + 1)]
        o.reweight_hansel_from_path(rec, path, 0.5)
        from collections import Counter
        cnt = Counter((i, j) for (_, _, i, j) in rec.calls if i != j)
        exp = Counter()
        for q in range(1, N):
            for p in range(0, q):
                exp[(p, q)] += 1
        for i in range(0, N - 1):
            exp[(i, i + 1)] += 1
        exp[(N - 1, N)] += 1
        exp[(N, N + 1)] += 1
        assert cnt == exp, N
        assert rec.calls[-1] == (path[N], '_', N, N + 1)
        if N == 4:
            assert [(i, j) for (_, _, i, j) in rec.calls] == [
                (0, 0), (0, 1), (0, 1), (1, 1), (1, 2), (0, 2), (1, 2), (2, 2), (2, 3),
                (0, 3), (1, 3), (2, 3), (3, 3), (3, 4), (4, 5)]


# ---- recovery control flow pinned to the reference's own code --------------------------------
# tests/golden/ref_recovery_*.npz were produced by tests/golden/make_golden.py, which imports
# /root/reference/gretel/gretel.py UNMODIFIED (generate_path, reweight_hansel_from_path) and
# drives it over OracleHansel.  The oracle's restatement of those loops must reproduce them.
import glob

import pytest

GOLDEN_RECOVERY = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_recovery_*.npz")))


@pytest.mark.parametrize("path", GOLDEN_RECOVERY, ids=[os.path.basename(p)[:-4] for p in GOLDEN_RECOVERY])
def test_oracle_reproduces_reference_recovery(path, c_oracle):
    z = np.load(path)
    N, L, v_site = int(z["N"]), int(z["L"]), str(z["v_site"])
    h = o.load_from_packed(z["rank"], z["off"], z["codes"], N, v_site=v_site)
    assert (h.n_slices, h.n_crumbs) == (int(z["n_slices"]), int(z["n_crumbs"]))
    assert np.array_equal(h.m, z["dense_before"])
    h.L = L
    its, _ = o.recover(h, N, max_paths=len(z["paths"]))
    assert len(its) == len(z["paths"]) and len(its) > 0
    for it, gp, gs in zip(its, z["paths"], z["stats"]):
        assert [o.CODE[c] for c in it["path"]] == list(gp)
        assert (it["hp_current"], it["hp_original"], it["min_marginal"], it["ratio"], it["removed"]) == tuple(gs)
    assert np.array_equal(h.m, z["dense_after"])
    # and the C twin, on the banded layout
    W = N + 1
    band, _ = c_oracle.ingest(z["rank"], z["off"], z["codes"], N, W)
    cur = band.astype(np.float32)
    orig = cur.copy()
    for gp, gs in zip(z["paths"], z["stats"]):
        pc, res = c_oracle.generate_path(cur, orig, N, W, L, v_site=v_site)
        assert list(pc) == list(gp) and res == tuple(gs[:3])
        assert c_oracle.reweight_path(cur, N, W, pc, gs[3]) == gs[4]
    assert np.array_equal(cur, o.band_of(h, W))
