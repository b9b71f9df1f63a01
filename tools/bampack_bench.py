"""Throughput of the BAM -> packed-reads producers (CPU only; run anywhere)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gretel_b200 import bamio
from tests.bamwriter import write_bam

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
rng = np.random.default_rng(1)
G = 100_000
starts = np.sort(rng.integers(0, G - 150, size=n))
bases = np.array(list("ACGT"))
seqs = bases[rng.integers(0, 4, size=(n, 150))]
t = time.time()
reads = [(0, int(s), 0, "r%d" % i, [("M", 150)], "".join(seqs[i])) for i, s in enumerate(starts)]
path = os.path.join(tempfile.mkdtemp(), "big.bam")
write_bam(path, [("ctg", G)], reads, block=60000, align=os.environ.get("BAM_ALIGN", "0") == "1")
print("wrote %d reads, %.1f MB BAM in %.1fs" % (n, os.path.getsize(path) / 1e6, time.time() - t))
snps = np.sort(rng.choice(np.arange(1, G + 1), size=10_000, replace=False))
vh = {"N": len(snps), "snp_rev": {i: int(p) for i, p in enumerate(snps)}}
for th in (1, 2, 4, 8, os.cpu_count()):
    t = time.time(); r = bamio.pack_bam_native(path, "ctg", 1, G, vh, n_threads=th); dt = time.time() - t
    print("native  %2d threads: %.3fs  %.2f M reads/s  (%d packed reads, %d codes)" % (th, dt, n / dt / 1e6, len(r[0]), len(r[2])))
sub = 20_000
write_bam(path + ".small", [("ctg", G)], reads[:sub])
t = time.time(); p = bamio.pack_bam(path + ".small", "ctg", 1, G, vh); dt = time.time() - t
print("python packer: %.3fs  %.3f M reads/s" % (dt, sub / dt / 1e6))
