"""Generates tests/golden/ref_recovery_*.npz by running the REFERENCE's own code.

Run in the build container only (it reads /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py

What is pinned: gretel/gretel.py's ``generate_path`` (:102-189) and
``reweight_hansel_from_path`` (:13-98) and the driver loop of gretel/cmd.py:148-161 are
imported from /root/reference UNMODIFIED and driven over ``OracleHansel`` objects (the
reference's un-vendored dependency ``hansel`` is stubbed with our oracle class; pysam/vcf
are stubbed because they are only needed by the BAM/VCF readers).  The outputs therefore
pin every piece of control flow of the recovery path to the real reference code: site
order, first-max tie-break, log10 accumulation, min-marginal, the reweight loop nest with
its double hit of adjacent pairs, the ratio clamp.  What stays unpinned is the arithmetic
inside Hansel itself (hanselx is not installable here) - see oracle/hansel_oracle.py.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hansel_oracle as o  # noqa: E402
from gretel_b200 import synth  # noqa: E402

# --- stubs for the reference's imports --------------------------------------------------
hansel_stub = types.ModuleType("hansel")
hansel_stub.Hansel = o.OracleHansel
sys.modules["hansel"] = hansel_stub
for name in ("pysam", "vcf"):
    sys.modules[name] = types.ModuleType(name)
sys.path.insert(0, "/root/reference")
import gretel.gretel as ref_gretel  # noqa: E402  (the reference, unmodified)

OUT = os.path.dirname(os.path.abspath(__file__))


def ref_recover(h, n_snps, max_paths):
    """gretel/cmd.py:79,148-161 with the reference's own functions."""
    orig = h.copy()
    its = []
    for _ in range(max_paths):
        path, prob, init_min = ref_gretel.generate_path(n_snps, h, orig)
        if path is None:
            break
        ratio = init_min
        if ratio < 0.01:            # MIN_REMOVE, cmd.py:157-160
            ratio = 0.01
        mag = ref_gretel.reweight_hansel_from_path(h, path, ratio)
        its.append((path, prob["hp_current"], prob["hp_original"], init_min, ratio, mag))
    return its


def case(seed, N, R, max_k, L, v_site):
    rng = np.random.default_rng(seed)
    rank, off, codes = synth.random_packed(rng, N, R, max_k, p_special=0.12)
    h = o.load_from_packed(rank, off, codes, N, v_site=v_site)
    h.L = L
    counts0 = h.m.copy()
    its = ref_recover(h, N, 6)
    return dict(rank=rank, off=off, codes=codes, N=N, L=L, v_site=v_site,
                n_slices=h.n_slices, n_crumbs=h.n_crumbs,
                dense_before=counts0, dense_after=h.m.copy(),
                paths=np.array([[o.CODE[s] for s in it[0]] for it in its], dtype=np.uint8).reshape(len(its), N + 1),
                stats=np.array([it[1:] for it in its], dtype=np.float64).reshape(len(its), 5))


def strain_case(seed, n_reads, L):
    w = synth.Workload("tiny", 1, 600, 40, n_reads, 120, 0.01, 0.005)
    d = synth.generate(w, seed=seed)
    N = w.n_snps
    h = o.load_from_packed(d["rank"], d["off"], d["codes"], N)
    h.L = L
    counts0 = h.m.copy()
    its = ref_recover(h, N, 5)
    return dict(rank=d["rank"], off=d["off"], codes=d["codes"], N=N, L=L, v_site="from",
                n_slices=h.n_slices, n_crumbs=h.n_crumbs, dense_before=counts0, dense_after=h.m.copy(),
                paths=np.array([[o.CODE[s] for s in it[0]] for it in its], dtype=np.uint8).reshape(len(its), N + 1),
                stats=np.array([it[1:] for it in its], dtype=np.float64).reshape(len(its), 5))


if __name__ == "__main__":
    import io, contextlib
    cases = {
        "ref_recovery_a": case(11, 9, 60, 6, 3, "from"),
        "ref_recovery_b": case(12, 14, 120, 8, 5, "to"),
        "ref_recovery_c": case(13, 6, 25, 4, 1, "from"),
        "ref_recovery_strains": strain_case(20260001, 400, 4),
    }
    for name, c in cases.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **c)
        print(name, "iterations:", len(c["paths"]), "N:", c["N"], "size:", os.path.getsize(os.path.join(OUT, name + ".npz")))
