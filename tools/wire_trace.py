"""Scratch: timeline of the dense-format chunks (HX_WIRE_TRACE=1)."""
import sys, os
os.environ["HX_WIRE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gretel_b200 import synth, util
from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
d = synth.generate(synth.WORKLOADS["metagenome"])
N, W = d["n_snps"], d["max_k"] - 1
for n_chunks in [int(x) for x in sys.argv[1].split(",")]:
    keep, chunks = [], []
    for c in util.dense_chunks(d["rank"], d["off"], d["codes"], n_chunks):
        pinned = torch.from_numpy(c.blob).pin_memory(); keep.append(pinned)
        chunks.append(c.rebased(pinned.numpy()))
    for it in range(3):
        h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
        for c in chunks:
            h.ingest_packed_dense(c, wait=False)
        if it == 2:
            sys.stderr.write("---- %d chunks\n" % n_chunks)
        else:
            os.environ.pop("X", None)
        h.ingest_totals(); h.finalize(); h.close()
