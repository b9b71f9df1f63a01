"""Scratch: device timeline (HX_WIRE_TRACE=1) and host time of the default hx_ingest_host path on config 3."""
import sys, os, time
os.environ["HX_WIRE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gretel_b200 import synth
from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
d = synth.generate(synth.WORKLOADS["metagenome"])
N, W = d["n_snps"], d["max_k"] - 1
pr = torch.from_numpy(d["rank"]).pin_memory(); po = torch.from_numpy(d["off"]).pin_memory(); pc = torch.from_numpy(d["codes"]).pin_memory()
for mode in (sys.argv[1:] or ["slim"]):
    os.environ["HX_HOST_PIPELINE"] = mode
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
        t1 = time.perf_counter()
        if it == 3: sys.stderr.write("---- %s\n" % mode)
        h.ingest_packed(pr.numpy(), po.numpy(), pc.numpy())
        t2 = time.perf_counter()
        h.finalize(); h.ingest_totals(); t3 = time.perf_counter(); h.close()
        torch.cuda.synchronize(); t4 = time.perf_counter()
        print("%s: create %.2f ingest_packed %.2f finalize %.2f close %.2f total %.2f ms" % (mode, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t4-t0)))
