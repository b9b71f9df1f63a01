"""gretel-snpper on the native BAM reader: aggressively call variants and print a placeholder VCF.

Mirrors gretel/snpper.py:12-56 of the reference (same flags, same output): per-position A,C,G,T counts
over every alignment (pysam.count_coverage(..., quality_threshold=0, read_callback='nofilter') there,
hx_count_coverage here), a site is reported when more than one base has a count above --depth.

    python -m gretel_b200.snpper --bam reads.bam --contig ctg [-s 1] [-e END] [--depth 0] > calls.vcf
"""
from __future__ import annotations

import argparse
import ctypes as C
import sys

import numpy as np

from . import _lib


def count_coverage(bam, contig, start0, end0, n_threads=1, device=0):
    """uint32 [4][end0-start0] counts of A,C,G,T (0-based half-open interval).  The BAM is decoded on the CPU, the
    per-position histogram runs on GPU ``device`` (hx_count_coverage_gpu); ``device=None`` counts on the CPU
    (hx_count_coverage: the checker of the GPU kernel in the tests, and what a CPU-only box can run)."""
    lib = _lib.load()
    out = np.zeros((4, max(0, end0 - start0)), dtype=np.uint32)
    if device is None:
        _lib.check(lib.hx_count_coverage(str(bam).encode(), str(contig).encode(), int(start0), int(end0),
                                         int(max(1, n_threads)), out.ctypes.data))
    else:
        _lib.check(lib.hx_count_coverage_gpu(str(bam).encode(), str(contig).encode(), int(start0), int(end0),
                                             int(max(1, n_threads)), int(device), out.ctypes.data))
    return out


def contig_length(bam, contig):
    lib = _lib.load()
    n = C.c_int32()
    _lib.check(lib.hx_bam_contig_length(str(bam).encode(), str(contig).encode(), C.byref(n)))
    return int(n.value)


def call_sites(counts, depth=0):
    """0-based offsets (into counts) of the sites with more than one base above ``depth`` (snpper.py:39-41)."""
    return np.nonzero((counts > depth).sum(axis=0) > 1)[0]


def main(argv=None, out=sys.stdout):
    p = argparse.ArgumentParser("Aggressively call for variants and generate a VCF",
                                epilog="NOTE: Coordinates are 1-based as they are for samtools")
    p.add_argument("--bam", required=True)
    p.add_argument("--contig", required=True)
    p.add_argument("-s", type=int, default=1)
    p.add_argument("-e", type=int)
    p.add_argument("--depth", type=int, default=0)
    p.add_argument("-@", "--threads", type=int, default=1)
    p.add_argument("--device", type=int, default=0, help="GPU that counts the bases (default 0)")
    p.add_argument("--cpu", action="store_true", help="count on the CPU instead")
    a = p.parse_args(argv)
    if not a.e:
        a.e = contig_length(a.bam, a.contig)
    s0 = a.s - 1
    counts = count_coverage(a.bam, a.contig, s0, a.e, n_threads=a.threads, device=None if a.cpu else a.device)
    out.write("##fileformat=VCFv4.2\n")
    for i in call_sites(counts, a.depth):
        out.write("\t".join([a.contig, str(int(i) + 1 + s0), ".", "A", "C,T,G", "0", ".", "INFO"]) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
