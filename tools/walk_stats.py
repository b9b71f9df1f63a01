"""How often the speculative blocks of the parallel-in-time walk are accepted (config-5 matrix, 50 haplotypes per L)."""
import ctypes as C, os, sys, time
sys.path.insert(0, ".")
import numpy as np
from gretel_b200 import synth, util, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d = synth.generate(synth.scaled(synth.WORKLOADS["metagenome"], n))
h = util.load_from_packed(d["rank"], d["off"], d["codes"], d["n_snps"], band_w=d["max_k"] - 1)
lib = _lib.load()
orig = h.copy()
for L in (1, 2, 4, 6, 8, 15):
    hc = orig.copy(); hc.L = L
    r0 = C.c_int(); lib.hx_debug_walk_redone(hc._h, C.byref(r0))
    t = time.perf_counter(); paths, stats = hc.recover_codes(orig, 50, 0.01); dt = time.perf_counter() - t
    r1 = C.c_int(); lib.hx_debug_walk_redone(hc._h, C.byref(r1))
    print("L=%2d: %d haplotypes in %.4f s; sites walked again %d of %d" % (L, len(paths), dt, r1.value - r0.value, d["n_snps"] * len(paths)), flush=True)
    hc.close()
