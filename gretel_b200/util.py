"""Ingestion entry points with the reference's names and argument meaning.

``load_from_bam`` / ``process_vcf`` / ``get_ref_len_from_bam`` mirror
gretel/util.py:33, :354, :10 of the reference.  BAM/VCF parsing stays on the CPU
(north_star) and produces the packed ``(rank, off, codes)`` arrays; the pair expansion
(util.py:226-286) runs in the CUDA ingestion kernel.
"""
from __future__ import annotations

import sys
from math import ceil

import numpy as np

from . import bamio
from .bamio import get_ref_len_from_bam, process_vcf  # noqa: F401  (re-exported, same names as the reference)
from .hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS


def band_width_for(off, floor=1):
    """Smallest band that holds every pair of every read: max SNPs on a read - 1."""
    off = np.asarray(off)
    if len(off) < 2:
        return max(1, floor)
    return max(int(np.diff(off).max()) - 1, floor, 1)


def compact_packed(off, codes):
    """(off int64[R+1], codes uint8) -> (klen uint16[R], codes4 uint8[ceil(n/2)], n_codes): the compact
    wire format of Hansel.ingest_packed_compact (two allele codes per byte, low nibble first)."""
    off = np.asarray(off, dtype=np.int64)
    k = np.diff(off)
    if len(k) and k.max() > 65535:
        raise ValueError("a read covers more than 65535 SNPs")
    c = np.ascontiguousarray(codes[off[0]:off[-1]] if len(off) else codes, dtype=np.uint8)
    n = len(c)
    if n & 1:
        c = np.concatenate([c, np.zeros(1, np.uint8)])
    codes4 = (c[0::2] | (c[1::2] << 4)).astype(np.uint8)
    return k.astype(np.uint16), codes4, n


def load_from_packed(rank, off, codes, n_snps, band_w=None, device=None, hansel=None, finalize=True,
                     quiet=True):
    """Packed reads -> Hansel (util.py:83 + 226-286 + 329-333).

    With ``finalize=False`` the integer counts stay pending so that partial matrices of
    several GPUs can be summed first (see gretel_b200.dist)."""
    if hansel is None:
        if band_w is None:
            band_w = band_width_for(off)
        hansel = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, n_snps, band_w=band_w, device=device)
    slices, crumbs, covered, _sent = hansel.ingest_packed(rank, off, codes)
    if finalize:
        hansel.finalize()
        set_totals(hansel, slices, crumbs, covered, quiet=quiet)
    return hansel


def set_totals(hansel, slices, crumbs, covered, quiet=True):
    """util.py:329-333."""
    hansel.n_slices = int(slices)
    hansel.n_crumbs = int(crumbs)
    if not quiet:
        sys.stderr.write("[NOTE] Loaded %d breadcrumbs from %d bread slices.\n" % (hansel.n_crumbs, hansel.n_slices))
    hansel.L = int(ceil(float(covered) / slices))        # ZeroDivisionError like the reference if no read has >=2 SNPs
    if not quiet:
        sys.stderr.write("[NOTE] Setting Gretel.L to %d\n" % hansel.L)
    return hansel


def load_from_bam(bam_path, target_contig, start_pos, end_pos, vcf_handler, use_end_sentinels=False,
                  n_threads=1, debug_reads=False, debug_pos=False, stepper="samtools", device=None,
                  band_w=None):
    """gretel/util.py:33.  Same signature and return value; ``n_threads`` is accepted for
    compatibility (the reference's window sharding is replaced by one GPU kernel and the
    result is independent of it, cf. tests/test_test.py:35).  ``use_end_sentinels`` is an
    experimental dead branch upstream (never passed, cmd.py:78) and is rejected here."""
    if use_end_sentinels:
        raise NotImplementedError("use_end_sentinels is never enabled by the reference (cmd.py:28,78)")
    # n_threads (the reference's number of BAM iterators, cmd.py:31) drives the native packer's threads
    rank, off, codes = bamio.pack_bam_native(bam_path, target_contig, start_pos, end_pos, vcf_handler,
                                             stepper=stepper, n_threads=n_threads)
    if band_w is None:
        # hold every ingested pair; at least N+1 for tiny regions so that the scalar API
        # (add/get_observation on arbitrary i<j) is band-resident like the reference's dense array
        band_w = band_width_for(off)
        if vcf_handler["N"] <= 256:
            band_w = max(band_w, vcf_handler["N"] + 1)
    return load_from_packed(rank, off, codes, vcf_handler["N"], band_w=band_w, device=device, quiet=False)
