"""Where the end-to-end ingestion time goes (scratch).  usage: e2e_breakdown.py [chunks]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gretel_b200 import synth, util
from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
arg = sys.argv[1] if len(sys.argv) > 1 else "4"
weights = [float(x) for x in arg.split(":")] if ":" in arg else None
n_chunks = len(weights) if weights else int(arg)
d = synth.generate(synth.WORKLOADS["metagenome"])
N, W = d["n_snps"], d["max_k"] - 1
keep, chunks = [], []
for c in util.dense_chunks(d["rank"], d["off"], d["codes"], n_chunks, weights=weights, native=True):
    pinned = torch.from_numpy(c.blob).pin_memory()
    keep.append(pinned)
    chunks.append(c.rebased(pinned.numpy()))
print("bytes", sum(c.nbytes for c in chunks))
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(6):
    t0 = T(); h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
    t1 = T()
    for c in chunks:
        h.ingest_packed_dense(c, wait=False)
    t1b = time.perf_counter()
    tot = h.ingest_totals()
    t2 = T(); h.finalize()
    t3 = T(); h.close()
    t4 = T()
    print("create %.2f  enqueue %.2f  wait %.2f  finalize %.2f  close %.2f  total %.2f ms" % (
        1e3*(t1-t0), 1e3*(t1b-t1), 1e3*(t2-t1b), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t4-t0)))
# copy alone / kernels alone
import ctypes
h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
for c in chunks: h.ingest_packed_dense(c, wait=False)
h.ingest_totals()
tot_bytes = sum(c.nbytes for c in chunks)
buf = torch.empty(tot_bytes, dtype=torch.uint8, device="cuda")
src = torch.empty(tot_bytes, dtype=torch.uint8).pin_memory()
for it in range(3):
    t0 = T(); buf.copy_(src, non_blocking=True); t1 = T()
    print("plain H2D of %d bytes: %.2f ms (%.1f GB/s)" % (tot_bytes, 1e3*(t1-t0), tot_bytes/(t1-t0)/1e9))
