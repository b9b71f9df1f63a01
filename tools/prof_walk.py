"""Scratch: one ingest + a few generate_path calls, for an ncu launch list of the recovery kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gretel_b200 import synth, util
from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
Ls = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 15]
w = synth.scaled(synth.WORKLOADS["metagenome"], n_reads)
d = synth.generate(w)
N, W = w.n_snps, d["max_k"] - 1
h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
tot = h.ingest_packed(d["rank"], d["off"], d["codes"])
h.finalize()
util.set_totals(h, tot[0], tot[1], tot[2])
o = h.copy()
for L in Ls:
    for rep in range(3):
        hh = h.copy(); hh.L = L
        r = hh.generate_path_codes(o)
        print("L=%d walk kernels %.3f ms" % (L, hh.kernel_ms("walk")))
