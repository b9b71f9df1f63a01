"""MMA-thread timers of k_l2_tiles (library built with HX_NVCC_DEFS=-DL2_PROFILE): python tools/lumma_prof.py ont 100000"""
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
    v = [x / (74.0 if os.environ.get("HX_LUMMA_PAIRS") == "1" else 148.0) for x in buf]
    if os.environ.get("HX_LUMMA_PAIRS") == "1":
        print("  (pairs: per leader; wait for the partner's stage %.0f clk)" % v[6])
    print("ingest %.3f ms; per CTA: MMA thread total %.0f clk = wait-full %.0f + issue %.0f + wait-acc-empty %.0f + rest %.0f; "
          "%.0f stages, %.0f MMAs (%.0f clk per MMA in the issue section)" % (
              ms, v[5], v[0], v[1], v[2], v[5] - v[0] - v[1] - v[2], v[3], v[4], v[1] / max(1.0, v[4])))
