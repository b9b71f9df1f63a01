"""Dependency-free BAM / VCF readers and the read packer (CPU side of the boundary).

north_star keeps parsing on the CPU: each alignment becomes a compact
``(first-SNP index, allele-code run)`` record.  The reference does this with a
pysam column pileup (gretel/util.py:137-210); here each alignment's CIGAR is walked
once against the sorted SNP-position array, which yields exactly the same per-read
allele run without the per-(read x column) Python objects:

* read key / mate separation ....... util.py:149-160 (each BAM record is its own read)
* window ownership / start clamp ... util.py:162-176 (union over windows == 1 thread)
* allele at a SNP column ........... util.py:180-190 (only the first character is
  ever used, util.py:238: '-' for a deletion, else the aligned base)
* rank ............................. util.py:198  (np.sum(region[1:LEFTMOST]) -> bisect)
* stepper filters .................. cmd.py:39,78 ("samtools": drop UNMAP/SECONDARY/
  QCFAIL/DUP and orphans; "all": the same without the orphan rule; "nofilter": none)

pysam is not required (it is not installed in the build image); BGZF is a series of
gzip members so the stdlib can read it.
"""
from __future__ import annotations

import gzip
import struct
from bisect import bisect_left, bisect_right

import numpy as np

SYMBOLS = ['A', 'C', 'G', 'T', 'N', '-', '_']       # util.py:83
UNSYMBOLS = ['N', '_']
_CODE = {s: i for i, s in enumerate(SYMBOLS)}
_SEQ_NT16 = "=ACMGRSVTWYHKDBN"
# any base that is not A/C/G/T is kept as its own character by the reference and
# would raise KeyError in Hansel; we map IUPAC ambiguity codes to 'N'.
_BASE2CODE = np.full(256, _CODE['N'], dtype=np.uint8)
for _s in "ACGT":
    _BASE2CODE[ord(_s)] = _CODE[_s]
    _BASE2CODE[ord(_s.lower())] = _CODE[_s]
_BASE2CODE[ord('-')] = _CODE['-']
_BASE2CODE[ord('_')] = _CODE['_']

BAM_FPAIRED, BAM_FPROPER_PAIR, BAM_FUNMAP = 0x1, 0x2, 0x4
BAM_FSECONDARY, BAM_FQCFAIL, BAM_FDUP = 0x100, 0x200, 0x400


class BamRecord:
    __slots__ = ("tid", "pos", "flag", "name", "cigar", "seq", "mapq")

    def __init__(self, tid, pos, flag, name, cigar, seq, mapq):
        self.tid, self.pos, self.flag, self.name = tid, pos, flag, name
        self.cigar, self.seq, self.mapq = cigar, seq, mapq

    @property
    def query_alignment_length(self):
        """pysam semantics: aligned query bases excluding soft clips (M/I/=/X)."""
        return sum(l for op, l in self.cigar if op in (0, 1, 7, 8))

    @property
    def reference_length(self):
        return sum(l for op, l in self.cigar if op in (0, 2, 3, 7, 8))


def read_bam(path):
    """Return (refs=[(name, length)], iterator of BamRecord) from an uncompressed-or-BGZF BAM."""
    with gzip.open(path, "rb") as fh:
        data = fh.read()
    if data[:4] != b"BAM\x01":
        raise ValueError("%s: not a BAM file" % path)
    p = 4
    (l_text,) = struct.unpack_from("<i", data, p); p += 4 + l_text
    (n_ref,) = struct.unpack_from("<i", data, p); p += 4
    refs = []
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", data, p); p += 4
        name = data[p:p + l_name - 1].decode(); p += l_name
        (l_ref,) = struct.unpack_from("<i", data, p); p += 4
        refs.append((name, l_ref))

    def records(p=p):
        n = len(data)
        while p + 4 <= n:
            (block_size,) = struct.unpack_from("<i", data, p); p += 4
            end = p + block_size
            (tid, pos, l_read_name, mapq, _bin, n_cigar, flag, l_seq,
             _ntid, _npos, _tlen) = struct.unpack_from("<iiBBHHHiiii", data, p)
            q = p + 32
            name = data[q:q + l_read_name - 1].decode(); q += l_read_name
            cigar = []
            for c in struct.unpack_from("<%dI" % n_cigar, data, q):
                cigar.append((c & 0xF, c >> 4))
            q += 4 * n_cigar
            packed = data[q:q + (l_seq + 1) // 2]
            seq = "".join(_SEQ_NT16[b >> 4] + _SEQ_NT16[b & 0xF] for b in packed)[:l_seq]
            yield BamRecord(tid, pos, flag, name, cigar, seq, mapq)
            p = end
    return refs, records()


def get_ref_len_from_bam(bam_path, target_contig):
    """gretel/util.py:10-31."""
    refs, _ = read_bam(bam_path)
    for name, ln in refs:
        if name == target_contig:
            return ln
    raise KeyError(target_contig)


def read_vcf_positions(vcf_path, contig_name):
    """1-based POS of every record on ``contig_name`` (no record-type filter, util.py:396-406)."""
    opener = gzip.open if str(vcf_path).endswith(".gz") else open
    out = []
    with opener(vcf_path, "rt") as fh:
        for line in fh:
            if not line or line[0] == "#":
                continue
            f = line.split("\t", 2)
            if len(f) < 2 or f[0] != contig_name:
                continue
            out.append(int(f[1]))
    return out


def process_vcf(vcf_path, contig_name, start_pos, end_pos):
    """gretel/util.py:354-414: N, snp_fwd (pos->idx), snp_rev (idx->pos), region mask."""
    region = np.zeros(end_pos + 1, dtype=int)
    snp_reverse, snp_forward = {}, {}
    i = 0
    for pos in read_vcf_positions(vcf_path, contig_name):
        if pos < start_pos or pos > end_pos:
            continue
        region[pos] = 1
        snp_reverse[i] = pos
        snp_forward[pos] = i
        i += 1
    return {"N": i, "snp_fwd": snp_forward, "snp_rev": snp_reverse, "region": region}


def _passes_stepper(flag, stepper):
    if stepper == "nofilter":
        return True
    if flag & (BAM_FUNMAP | BAM_FSECONDARY | BAM_FQCFAIL | BAM_FDUP):
        return False
    if stepper == "samtools" and (flag & BAM_FPAIRED) and not (flag & BAM_FPROPER_PAIR):
        return False                                      # orphan rule
    return True


def alleles_for_record(rec, snp_pos_sorted, start_pos, end_pos):
    """(rank, [allele chars]) for one alignment, or None if it is skipped.

    Equivalent to what util.py:137-209 accumulates for this read over all pileup
    columns (the union over work blocks; see module docstring)."""
    leftmost = rec.pos + 1                                  # util.py:162
    if leftmost < start_pos:                                # util.py:165-171
        if rec.pos + 1 + rec.query_alignment_length < start_pos:
            return None
        leftmost = start_pos
    rank = bisect_left(snp_pos_sorted, leftmost)            # == np.sum(region[1:leftmost])
    ref_end_1 = rec.pos + rec.reference_length              # last covered 1-based position
    last = min(ref_end_1, end_pos)
    hi = bisect_right(snp_pos_sorted, last)
    if hi <= rank:
        return rank, []
    alleles = []
    # walk CIGAR once; cursor over the wanted SNP positions
    want = snp_pos_sorted[rank:hi]
    wi = 0
    rpos = rec.pos + 1                                      # 1-based ref cursor
    qpos = 0
    for op, ln in rec.cigar:
        if wi >= len(want):
            break
        if op in (0, 7, 8):                                 # M = X
            while wi < len(want) and want[wi] < rpos + ln:
                if want[wi] >= rpos:
                    alleles.append(rec.seq[qpos + (want[wi] - rpos)])
                wi += 1
            rpos += ln; qpos += ln
        elif op in (2, 3):                                  # D / N -> '-' (is_del)
            while wi < len(want) and want[wi] < rpos + ln:
                if want[wi] >= rpos:
                    alleles.append('-')
                wi += 1
            rpos += ln
        elif op in (1, 4):                                  # I / S consume query only
            qpos += ln
        # H, P consume nothing
    return rank, alleles


PYSAM_MAX_DEPTH = 8000      # pysam's AlignmentFile.pileup(max_depth=8000); gretel never changes it (util.py:137)


def snp_positions(vcf_handler):
    """Sorted UNIQUE 1-based SNP positions of a process_vcf() result.  The reference marks ``region[pos]`` once
    per position (util.py:402), so a VCF that repeats a POS (split multi-allelic records) yields one pileup column
    and one rank step for it, while ``N`` still counts every record (util.py:404-406)."""
    return sorted(set(vcf_handler["snp_rev"][i] for i in range(vcf_handler["N"])))


class _DepthCap:
    """htslib's bam_plp buffer limit as pysam's pileup applies it: a read that is not the first of its start
    position is dropped while ``max_depth`` admitted reads are still buffered (end > the previous column)."""

    def __init__(self, max_depth):
        import heapq
        self.max_depth, self.ends, self.pos, self.hq = max_depth, [], None, heapq

    def admit(self, rec):
        if not self.max_depth:
            return True
        beg, end = rec.pos, rec.pos + max(1, rec.reference_length)
        if beg != self.pos:
            while self.ends and self.ends[0] <= beg - 1:
                self.hq.heappop(self.ends)
            self.pos = beg
        elif len(self.ends) >= self.max_depth:
            return False
        self.hq.heappush(self.ends, end)
        return True


def pack_reads(records, snp_pos_sorted, start_pos, end_pos, target_tid, stepper="samtools", max_depth=0):
    """Pack alignments into ``(rank int32[R], off int64[R+1], codes uint8[sum k])``.

    Reads with fewer than two covered SNPs carry no pair evidence (util.py:230) and
    are dropped here; the order of the survivors is BAM order (coordinate-sorted BAM
    => non-decreasing rank)."""
    ranks, offs, chunks = [], [0], []
    total = 0
    cap = _DepthCap(max_depth)
    for rec in records:
        if rec.tid != target_tid or not _passes_stepper(rec.flag, stepper):
            continue
        if rec.pos >= end_pos or rec.pos + max(1, rec.reference_length) <= start_pos - 1:
            continue                                        # not fetched by pileup(start-1, end)
        if not cap.admit(rec):
            continue
        if rec.pos + 1 > end_pos:
            continue
        res = alleles_for_record(rec, snp_pos_sorted, start_pos, end_pos)
        if res is None:
            continue
        rank, alleles = res
        if len(alleles) < 2:
            continue
        ranks.append(rank)
        chunks.append(_BASE2CODE[np.frombuffer("".join(a[0] for a in alleles).encode(), dtype=np.uint8)])
        total += len(alleles)
        offs.append(total)
    codes = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
    return (np.asarray(ranks, dtype=np.int32), np.asarray(offs, dtype=np.int64),
            np.ascontiguousarray(codes, dtype=np.uint8))


def pack_bam_native(bam_path, target_contig, start_pos, end_pos, vcf_handler, stepper="samtools", n_threads=1,
                    max_depth=0, stages=None):
    """BAM + process_vcf() output -> packed reads, by the multi-threaded streaming C++ packer of libhanselx.so
    (hx_pack_bam_ex: BGZF inflate + one CIGAR walk per alignment).  Same result as pack_bam().  ``stages``: a
    dict that receives the seconds spent per stage (read, inflate, scan, depth, walk, gather)."""
    import ctypes as C

    from . import _lib
    lib = _lib.load()
    snp_pos = np.ascontiguousarray(snp_positions(vcf_handler), dtype=np.int32)
    out = _lib.HxPacked()
    secs = (C.c_double * 6)()
    rc = lib.hx_pack_bam_ex(str(bam_path).encode(), str(target_contig).encode(), int(start_pos), int(end_pos),
                            snp_pos.ctypes.data, len(snp_pos), {"samtools": 0, "all": 1, "nofilter": 2}[stepper],
                            int(max(1, n_threads)), int(max_depth or 0), C.byref(out), secs)
    _lib.check(rc)
    if stages is not None:
        stages.update(dict(zip(("read", "inflate", "scan", "depth", "walk", "gather"), (float(x) for x in secs))))
        stages["records"] = int(out.n_records)
    try:
        R, n = int(out.n_reads), int(out.n_codes)
        rank = np.ctypeslib.as_array(out.rank, shape=(max(R, 1),))[:R].copy()
        off = np.ctypeslib.as_array(out.off, shape=(R + 1,)).copy()
        codes = np.ctypeslib.as_array(out.codes, shape=(max(n, 1),))[:n].copy()
    finally:
        lib.hx_pack_free(C.byref(out))
    return rank, off, codes


def pack_bam(bam_path, target_contig, start_pos, end_pos, vcf_handler, stepper="samtools", max_depth=0):
    """BAM + process_vcf() output -> packed reads (dependency-free Python reader)."""
    refs, recs = read_bam(bam_path)
    names = [n for n, _ in refs]
    tid = names.index(target_contig)
    return pack_reads(recs, snp_positions(vcf_handler), start_pos, end_pos, tid, stepper=stepper, max_depth=max_depth)
