"""Quick parity + timing check of the long-read tensor-core ingestion kernel (kernel 7) against the C oracle."""
import os, sys, time
os.environ["HX_HOST_PIPELINE"] = "off"      # one launch per ingest_packed: kernel_ms is the whole ingestion
import numpy as np
sys.path.insert(0, ".")
from gretel_b200 import synth
from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
from oracle import c_oracle


def run(rank, off, codes, N, W, kernel):
    """-> band after one ingestion into a cleared matrix, band after a second one on top, totals, ms of each."""
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
    h.set_ingest_kernel(kernel)
    h.ingest_packed(rank, off, codes)                 # warm-up (scratch allocations)
    h.reset_counts()
    t = h.ingest_packed(rank, off, codes)
    ms1 = h.kernel_ms("ingest")
    cnt1 = h.counts_to_host() if hasattr(h, "counts_to_host") else None
    h.ingest_packed(rank, off, codes)
    ms2 = h.kernel_ms("ingest")
    b2 = h.band()
    h.close()
    return b2, t, ms1, ms2


def check(name, rank, off, codes, N, W, kernels=(7, 3), reps=1):
    ref, rt = c_oracle.ingest(rank, off, codes, N, W)
    ref = ref.astype(np.float32) * 2
    for kernel in kernels:
        band, tot, ms1, ms2 = run(rank, off, codes, N, W, kernel)
        ok = np.array_equal(band, ref) and tot == tuple(int(x) for x in rt)
        nbad = int((band != ref).sum())
        print("%-28s kernel %d: %s  bad cells %d  totals %s vs %s  cleared %.3f ms, on top %.3f ms" % (
            name, kernel, "OK" if ok else "MISMATCH", nbad, tot, tuple(int(x) for x in rt), ms1, ms2), flush=True)
        if not ok and nbad:
            idx = np.argwhere(band != ref)[:12]
            for i in idx:
                print("    pj=%d d=%d a=%d b=%d got %g want %g" % (i[0], i[1] + 1, i[2], i[3], band[tuple(i)], ref[tuple(i)]))


rng = np.random.default_rng(1)
k = np.full(64, 4); off = np.concatenate([[0], np.cumsum(k)]).astype(np.int64)
check("tiny-1rank", np.zeros(64, np.int32) + 2, off, rng.integers(0, 4, size=off[-1]).astype(np.uint8), 12, 3)
r, o, c = synth.random_packed(rng, 60, 3000, 12, p_special=0.0)
check("random k<=12 no specials", r, o, c, 60, 15)
r, o, c = synth.random_packed(rng, 150, 5000, 30, p_special=0.1)
check("random k<=30 specials", r, o, c, 150, 29)
r, o, c = synth.random_packed(rng, 700, 20000, 300, p_special=0.1, sort=False)
check("random k<=300 unsorted", r, o, c, 700, 320)
full = len(sys.argv) > 1
for name, n in (("ont", 3000), ("mid", 50_000), ("ont", 100_000 if full else 20_000), ("mid", 1_000_000 if full else 100_000)):
    d = synth.generate(synth.scaled(synth.WORKLOADS[name], n))
    check("%s %d" % (name, n), d["rank"], d["off"], d["codes"], d["n_snps"], d["max_k"] - 1, reps=3)
