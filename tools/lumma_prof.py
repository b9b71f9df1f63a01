"""Role timers of k_l2_tiles (library built with HX_NVCC_DEFS=-DL2_PROFILE): python tools/lumma_prof.py ont 100000"""
import sys, os, ctypes as C
os.environ["HX_HOST_PIPELINE"] = "off"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gretel_b200 import synth, _lib
from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
name = sys.argv[1] if len(sys.argv) > 1 else "ont"
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
d = synth.generate(synth.scaled(synth.WORKLOADS[name], n_reads))
h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, d["n_snps"], band_w=d["max_k"] - 1)
h.set_ingest_kernel(7)
L = _lib.load()
buf = (C.c_ulonglong * 16)()
for it in range(3):
    h.reset_counts()
    L.hx_debug_l2_prof(buf, 1)
    h.ingest_packed(d["rank"], d["off"], d["codes"])
    ms = h.kernel_ms("ingest")
    L.hx_debug_l2_prof(buf, 0)
    v = list(buf)
    n = 148.0
    print("ingest %.3f ms" % ms)
    print("  producer: wait-empty %.0f clk/CTA over %.0f stages (%.0f clk each); tile setup %.0f clk/CTA over %.0f tiles (%.0f each); total %.0f clk" % (
        v[0] / n, v[1] / n, v[0] / max(1, v[1]), v[2] / n, v[3] / n, v[2] / max(1, v[3]), v[12] / n))
    print("  mma: wait-full %.0f clk/CTA over %.0f stages (%.0f each); wait-acc-empty %.0f clk/CTA over %.0f tiles (%.0f each)" % (
        v[4] / n, v[5] / n, v[4] / max(1, v[5]), v[6] / n, v[7] / n, v[6] / max(1, v[7])))
    print("  mma: fence+descriptors %.0f clk/CTA, MMA issue %.0f, commit %.0f" % (v[13] / n, v[14] / n, v[15] / n))
    print("  epilogue warp 2: wait-acc-full %.0f clk/CTA over %.0f tiles (%.0f each)" % (v[8] / n, v[9] / n, v[8] / max(1, v[9])))
