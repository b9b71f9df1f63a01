"""Per-role cycle accounting of the tensor-core ingestion kernel (build with HX_NVCC_DEFS=-DUM_PROFILE)."""
import ctypes as C, os, sys
os.environ["HX_HOST_PIPELINE"] = "off"
import numpy as np
sys.path.insert(0, ".")
from gretel_b200 import synth, _lib
from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = synth.generate(synth.scaled(synth.WORKLOADS["metagenome"], n))
lib = _lib.load()
h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, d["n_snps"], band_w=d["max_k"] - 1)
h.set_ingest_kernel(6)
out = (C.c_ulonglong * 32)()
for it in range(3):
    h.reset_counts() if it else None
    lib.hx_debug_um_prof(out, 1)
    h.ingest_packed(d["rank"], d["off"], d["codes"])
    ms = h.kernel_ms("ingest")
    lib.hx_debug_um_prof(out, 0)
v = np.array(list(out), dtype=np.float64)
ctas = 148
print("kernel %.3f ms = %.0f cycles @1.965GHz" % (ms, ms * 1.965e6))
jobs = v[5]
print("expander: jobs/CTA %.0f; per job cycles: walk+off %.0f, issue codes loads %.0f, stage+slot wait %.0f, expand+mma %.0f, side %.0f; per warp total %.0f" % (
    jobs / ctas, v[0] / jobs, v[1] / jobs, v[2] / jobs, v[3] / jobs, v[4] / jobs, v[:5].sum() / ctas / 24))
r = v[19]
print("readout: runs/CTA %.0f; per run: flush %.0f, wait acc %.0f, read+add+zero+barrier %.0f; total/CTA %.0f" % (r / ctas, v[16] / r, v[17] / r, v[18] / r, v[16:19].sum() / ctas))
nw = ctas * 32
print("per warp avg cycles: prologue %.0f, roles %.0f, epilogue %.0f; max CTA total %.0f; max expander roles %.0f, max readout roles %.0f" % (
    v[24] / nw, v[25] / nw, v[26] / nw, v[27], v[28], v[29]))
