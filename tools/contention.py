"""Scratch: does a concurrent host->device copy slow the pair-expansion kernel down?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gretel_b200 import synth, util
from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
full = synth.generate(synth.WORKLOADS["metagenome"])
N, W = full["n_snps"], full["max_k"] - 1
R = len(full["rank"])
dev = torch.device("cuda", 0)
t_rank = torch.from_numpy(full["rank"]).to(dev); t_off = torch.from_numpy(full["off"]).to(dev); t_codes = torch.from_numpy(full["codes"]).to(dev)
h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
h.counts_buffer()
src = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
dst = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
dsrc = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cs = torch.cuda.Stream()
def run(label, bg):
    ms = []
    for it in range(5):
        h.reset_counts()
        torch.cuda.synchronize()
        if bg == "h2d":
            with torch.cuda.stream(cs):
                for _ in range(2): dst.copy_(src, non_blocking=True)
        elif bg == "d2d":
            with torch.cuda.stream(cs):
                for _ in range(40): dst.copy_(dsrc, non_blocking=True)
        h.ingest_device(t_rank.data_ptr(), t_off.data_ptr(), t_codes.data_ptr(), R)
        h.ingest_totals()
        torch.cuda.synchronize()
        ms.append(h.kernel_ms("ingest"))
    print("%-28s kernel %.3f ms (min) %.3f (median)" % (label, min(ms), sorted(ms)[2]))
run("alone", None)
run("with H2D copy in flight", "h2d")
run("with D2D copies in flight", "d2d")
run("alone again", None)

# quarter-sized launches back to back, alone and with a copy in flight
cuts = [0] + [int(np.searchsorted(full["off"], full["off"][-1] * q // 4)) for q in (1, 2, 3)] + [R]
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(4)]
st = torch.cuda.ExternalStream(h.stream, device=dev)
def run4(label, bg):
    best = None
    for it in range(5):
        h.reset_counts(); torch.cuda.synchronize()
        if bg:
            with torch.cuda.stream(cs):
                for _ in range(2): dst.copy_(src, non_blocking=True)
        for q in range(4):
            a, b = cuts[q], cuts[q + 1]
            ev[q][0].record(st)
            h.ingest_device(t_rank.data_ptr() + 4 * a, t_off.data_ptr() + 8 * a, t_codes.data_ptr(), b - a)
            ev[q][1].record(st)
        h.ingest_totals(); torch.cuda.synchronize()
        ms = [x.elapsed_time(y) for x, y in ev]
        if best is None or sum(ms) < sum(best): best = ms
    print("%-28s quarters %s" % (label, " ".join("%.3f" % m for m in best)))
run4("quarters alone", False)
run4("quarters with H2D in flight", True)
run4("quarters alone again", False)
