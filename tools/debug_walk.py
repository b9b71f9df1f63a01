"""Scratch: locate the first site where the GPU walk and the C oracle disagree and show the weights there."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gretel_b200 import synth, util
from oracle import c_oracle, hansel_oracle as o

L = int(sys.argv[1]) if len(sys.argv) > 1 else 9
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4200 + L
c_oracle.build()
rng = np.random.default_rng(seed)
N = 150
rank, off, codes = synth.random_packed(rng, N, 900, 38, p_special=0.05)
W = int(np.diff(off).max()) - 1
band, _ = c_oracle.ingest(rank, off, codes, N, W)
h = util.load_from_packed(rank, off, codes, N, band_w=W)
h.L = L
cur = band.astype(np.float32)
pc, res = c_oracle.generate_path(cur, cur.copy(), N, W, L)
r = h.generate_path_codes(h.copy())
print("W", W, "oracle", res, "gpu", r[1:])
d = np.nonzero(np.asarray(r[0]) != np.asarray(pc))[0]
print("differing sites", d[:10])
if len(d):
    s0 = int(d[0])
    ho = o.load_from_packed(rank, off, codes, N)
    ho.L = L
    syms = "ACGTN-_"
    for name, path in (("oracle", pc), ("gpu", r[0])):
        p = [syms[c] for c in path[:s0]]
        ew = ho.get_edge_weights_at(s0, p)
        print(name, "prefix tail", p[-10:], "chose", syms[path[s0]])
        print("   weights at site", s0, {k: "%.17g" % v for k, v in ew.items()})
