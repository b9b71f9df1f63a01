"""Ad-hoc timing on the GPU box (scratch; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gretel_b200 import synth, util
from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS

name = sys.argv[1] if len(sys.argv) > 1 else "metagenome"
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
kernel = int(sys.argv[3]) if len(sys.argv) > 3 else 0
w = synth.scaled(synth.WORKLOADS[name], n_reads)
t = time.time(); d = synth.generate(w); print("gen %.1fs reads=%d max_k=%d pairs=%d" % (time.time() - t, len(d["rank"]), d["max_k"], d["n_pairs"]))
N, W = w.n_snps, d["max_k"] - 1
h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
h.set_ingest_kernel(kernel)
for it in range(4):
    t = time.time()
    tot = h.ingest_packed(d["rank"], d["off"], d["codes"])
    dt = time.time() - t
    ms = h.kernel_ms("ingest")
    print("ingest it%d e2e %.1f ms kernel %.3f ms -> %.2f Gobs/s (kernel)" % (it, dt * 1e3, ms, d["n_pairs"] / ms / 1e6), tot)
h.finalize()
util.set_totals(h, tot[0] // 4, tot[1] // 4, tot[2] // 4)
o = h.copy()
for L in (1, h.L, 8):
    hh = h.copy(); hh.L = L
    t = time.time(); r = hh.generate_path_codes(o); dt = time.time() - t
    print("L=%d generate_path %.2f ms (walk kernels %.2f ms) -> %.1f us/site" % (L, dt * 1e3, hh.kernel_ms("walk"), hh.kernel_ms("walk") * 1e3 / N), None if r[0] is None else r[1:])
    if r[0] is not None:
        t = time.time(); rem = hh.reweight_path_codes(r[0], max(r[3], 0.01)); dt = time.time() - t
        print("   reweight %.2f ms (kernel %.3f ms) removed %.1f" % (dt * 1e3, hh.kernel_ms("reweight"), rem))
    t = time.time(); p, s = hh.recover_codes(o, 10); dt = time.time() - t
    print("   recover 10 resident: %.1f ms, found %d" % (dt * 1e3, len(p)))
