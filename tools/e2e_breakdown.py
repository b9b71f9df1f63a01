"""Where the end-to-end ingestion time goes (scratch)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gretel_b200 import synth, util
from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
d = synth.generate(synth.WORKLOADS["metagenome"])
N, W = d["n_snps"], d["max_k"] - 1
klen, codes4, n_codes = util.compact_packed(d["off"], d["codes"])
pr = torch.from_numpy(d["rank"]).pin_memory().numpy(); pk = torch.from_numpy(klen).pin_memory().numpy(); pc = torch.from_numpy(codes4).pin_memory().numpy()
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(5):
    t0 = T(); h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
    t1 = T(); tot = h.ingest_packed_compact(pr, pk, pc, n_codes)
    t2 = T(); h.finalize()
    t3 = T(); h.close()
    t4 = T()
    print("create %.2f  ingest(h2d+kernels+totals) %.2f [kernel %.2f]  finalize %.2f  close %.2f  total %.2f ms" % (
        1e3*(t1-t0), 1e3*(t2-t1), h.kernel_ms("ingest") if False else -1, 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t4-t0)))
