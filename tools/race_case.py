"""Small ingestion + recovery run for compute-sanitizer racecheck/memcheck (scratch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gretel_b200 import synth, util, gretel
from oracle import c_oracle
for name, n in (("hiv", 6000), ("metagenome", 40000), ("ont", 200)):
    w = synth.scaled(synth.WORKLOADS[name], n)
    d = synth.generate(w)
    W = d["max_k"] - 1
    h = util.load_from_packed(d["rank"], d["off"], d["codes"], w.n_snps, band_w=W)
    ref, rt = c_oracle.ingest(d["rank"], d["off"], d["codes"], w.n_snps, W)
    assert np.array_equal(h.band(), ref.astype(np.float32)), name
    h.L = min(h.L, 12)
    its, _ = gretel.recover(h, w.n_snps, max_paths=2)
    print(name, "ok", h.n_crumbs, len(its))
