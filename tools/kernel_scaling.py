"""Scratch: ingestion kernel time vs number of reads (same 10k-SNP region), to expose fixed per-launch cost."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gretel_b200 import synth, util
from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
full = synth.generate(synth.WORKLOADS["metagenome"])
N, W = full["n_snps"], full["max_k"] - 1
R = len(full["rank"])
h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
for frac, mode in [(1, "all"), (2, "prefix"), (4, "prefix"), (8, "prefix"), (2, "thin"), (4, "thin"), (8, "thin"), (20, "thin")]:
    if mode == "prefix":          # a contiguous chunk of the sorted reads (what a chunked ingestion sees)
        sel = slice(0, R // frac)
        rank = full["rank"][sel]; off = full["off"][:R // frac + 1]; codes = full["codes"]
    elif mode == "thin":          # every frac-th read (lower coverage over the whole region)
        idx = np.arange(0, R, frac)
        k = np.diff(full["off"])[idx]
        off = np.zeros(len(idx) + 1, np.int64); np.cumsum(k, out=off[1:])
        src = np.repeat(full["off"][:-1][idx], k) + (np.arange(off[-1]) - np.repeat(off[:-1], k))
        rank = full["rank"][idx]; codes = full["codes"][src]
    else:
        rank, off, codes = full["rank"], full["off"], full["codes"]
    ms = []
    for it in range(4):
        h.reset_counts()
        h.ingest_packed(rank, off, codes)
        ms.append(h.kernel_ms("ingest"))
    print("%-6s 1/%-2d reads %8d  kernel %.3f ms  (%.1f ns per 1k reads)" % (mode, frac, len(rank), min(ms), 1e6 * min(ms) / len(rank) * 1e3 / 1e3))
