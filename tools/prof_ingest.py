"""One ingestion launch for ncu (scratch)."""
import sys, os
os.environ["HX_HOST_PIPELINE"] = "off"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gretel_b200 import synth
from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
name = sys.argv[1] if len(sys.argv) > 1 else "metagenome"
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
kernel = int(sys.argv[3]) if len(sys.argv) > 3 else 0
d = synth.generate(synth.scaled(synth.WORKLOADS[name], n_reads))
h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, d["n_snps"], band_w=d["max_k"] - 1)
h.set_ingest_kernel(kernel)
for _ in range(3):
    print(h.ingest_packed(d["rank"], d["off"], d["codes"]), h.kernel_ms("ingest"))
