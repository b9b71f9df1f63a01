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
# dense wire format (decode scans, unpack, exception patch; chunked, three streams) + fixed-point walk depths
from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
for name, n in (("metagenome", 40000), ("ont", 200)):
    w = synth.scaled(synth.WORKLOADS[name], n)
    d = synth.generate(w)
    W = d["max_k"] - 1
    ref, rt = c_oracle.ingest(d["rank"], d["off"], d["codes"], w.n_snps, W)
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, w.n_snps, band_w=W)
    for c in util.dense_chunks(d["rank"], d["off"], d["codes"], 3):
        h.ingest_packed_dense(c, wait=False)
    tot = h.ingest_totals()
    assert tot == tuple(int(x) for x in rt), name
    assert np.array_equal(h.band(), ref.astype(np.float32)), name
    util.set_totals(h, tot[0], tot[1], tot[2])
    for L in (1, 3, 9, 17, 30):
        hh = h.copy(); hh.L = L
        r = hh.generate_path_codes(h)
        pc, res = c_oracle.generate_path(ref.astype(np.float32), ref.astype(np.float32), w.n_snps, W, L)
        assert (pc is None) == (r[0] is None), (name, L)
    print(name, "dense + walk ok", tot)
