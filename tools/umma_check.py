"""Quick parity + timing check of the tensor-core ingestion kernel (kernel 6) against the C oracle."""
import os, sys, time
os.environ["HX_HOST_PIPELINE"] = "off"      # one launch per ingest_packed: kernel_ms is the whole ingestion
import numpy as np
sys.path.insert(0, ".")
from gretel_b200 import synth
from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
from oracle import c_oracle


def run(rank, off, codes, N, W, kernel):
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
    h.set_ingest_kernel(kernel)
    t = h.ingest_packed(rank, off, codes)
    ms = h.kernel_ms("ingest")
    b = h.band()
    h.close()
    return b, t, ms


def check(name, rank, off, codes, N, W):
    ref, rt = c_oracle.ingest(rank, off, codes, N, W)
    ref = ref.astype(np.float32)
    for kernel in (6, 2):
        band, tot, ms = run(rank, off, codes, N, W, kernel)
        ok = np.array_equal(band, ref) and tot == tuple(int(x) for x in rt)
        nbad = int((band != ref).sum())
        print("%-28s kernel %d: %s  bad cells %d  totals %s vs %s  %.3f ms" % (
            name, kernel, "OK" if ok else "MISMATCH", nbad, tot, tuple(int(x) for x in rt), ms), flush=True)
        if not ok and nbad:
            idx = np.argwhere(band != ref)[:8]
            for i in idx:
                print("    pj=%d d=%d a=%d b=%d got %g want %g" % (i[0], i[1] + 1, i[2], i[3], band[tuple(i)], ref[tuple(i)]))


rng = np.random.default_rng(1)
# one rank, 64 reads of 4 SNPs, no specials
k = np.full(64, 4); off = np.concatenate([[0], np.cumsum(k)]).astype(np.int64)
check("tiny-1rank", np.zeros(64, np.int32) + 2, off, rng.integers(0, 4, size=off[-1]).astype(np.uint8), 12, 3)
r, o, c = synth.random_packed(rng, 60, 3000, 12, p_special=0.0)
check("random k<=12 no specials", r, o, c, 60, 15)
r, o, c = synth.random_packed(rng, 150, 5000, 30, p_special=0.1)
check("random k<=30 specials", r, o, c, 150, 29)
for n in (200_000, 10_000_000 if len(sys.argv) > 1 else 1_000_000):
    d = synth.generate(synth.scaled(synth.WORKLOADS["metagenome"], n))
    check("metagenome %d" % n, d["rank"], d["off"], d["codes"], d["n_snps"], d["max_k"] - 1)
