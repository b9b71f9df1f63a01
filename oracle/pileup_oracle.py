"""TEST INFRASTRUCTURE ONLY - a literal, column-by-column restatement of the reference's read extraction,
gretel/util.py:112-210 (``bam_worker`` up to the per-read dictionary) and :288-301 (the work blocks), over
already-decoded alignment records.  The product packers (hx_pack_bam, gretel_b200.bamio.pack_reads) walk every
alignment's CIGAR once instead; tests/test_pileup_oracle.py holds them to this file.

The reference reads the BAM through ``pysam.AlignmentFile.pileup`` (util.py:137).  pysam / htslib are not
installable here, so the pileup engine itself is restated from its documented behaviour (htslib ``bam_plp``;
this part is a recollection of htslib, PARITY UNPINNED by any reference fixture beyond tests/data/test.bam):

* ``pileup(reference, start, stop, stepper=...)`` fetches, in file order, the alignments of ``reference`` that
  overlap the 0-based half-open interval [start, stop) and pass the stepper ("all": drop UNMAP / SECONDARY /
  QCFAIL / DUP; "samtools": additionally drop paired reads that are not properly paired; "nofilter": none);
* ``truncate=False``: every column covered by a fetched read is yielded, also outside [start, stop);
* a column lists the fetched reads covering it in file order; inside a deletion (D) or a reference skip (N) the
  read is present with ``is_del`` set; insertions, soft clips, hard clips and pads consume no reference;
* ``min_base_quality=0`` keeps every base (util.py:137; CHANGELOG.rst:26-30);
* ``max_depth`` (pysam default 8000, never changed by gretel): the engine buffers reads; a read that is not the
  first one of its start position is dropped while ``max_depth`` reads are still buffered (a read leaves the
  buffer once the column before the new start position has been emitted and lies at or past its end).

A record is ``(tid, pos0, flag, name, cigar, seq)`` with ``cigar = [(op, length)]``, ops as in BAM
(0 M, 1 I, 2 D, 3 N, 4 S, 5 H, 6 P, 7 =, 8 X).
"""
from __future__ import annotations

import heapq

import numpy as np

FPAIRED, FPROPER, FUNMAP, FREAD1, FREAD2 = 0x1, 0x2, 0x4, 0x40, 0x80
FSECONDARY, FQCFAIL, FDUP = 0x100, 0x200, 0x400

CODE = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4, "-": 5, "_": 6}


def _ref_len(cigar):
    return sum(l for op, l in cigar if op in (0, 2, 3, 7, 8))


def _query_alignment_length(cigar):          # pysam: aligned query bases, soft clips excluded
    return sum(l for op, l in cigar if op in (0, 1, 7, 8))


def _stepper_ok(flag, stepper):
    if stepper == "nofilter":
        return True
    if flag & (FUNMAP | FSECONDARY | FQCFAIL | FDUP):
        return False
    if stepper == "samtools" and (flag & FPAIRED) and not (flag & FPROPER):
        return False
    return True


def _at(rec, refpos0):
    """(query_position, is_del) of the alignment at a 0-based reference position it covers."""
    tid, pos, flag, name, cigar, seq = rec
    r, q = pos, 0
    for op, ln in cigar:
        if op in (0, 7, 8):
            if r <= refpos0 < r + ln:
                return q + (refpos0 - r), False
            r += ln
            q += ln
        elif op in (2, 3):
            if r <= refpos0 < r + ln:
                return q, True
            r += ln
        elif op in (1, 4):
            q += ln
    raise AssertionError("position not covered")


def pileup(records, tid, start0, stop0, stepper="samtools", max_depth=8000):
    """Yield (reference_pos0, [record index, ...]) like ``bam.pileup(reference, start0, stop0, stepper=...)``."""
    fetched = []
    live, cur = [], None
    for i, rec in enumerate(records):
        rtid, pos, flag, name, cigar, seq = rec
        if rtid != tid or pos < 0 or not _stepper_ok(flag, stepper):
            continue
        end = pos + max(1, _ref_len(cigar))
        if pos >= stop0 or end <= start0:
            continue
        if pos != cur:
            while live and live[0] <= pos - 1:
                heapq.heappop(live)
            cur = pos
        elif max_depth and len(live) >= max_depth:
            continue
        heapq.heappush(live, end)
        fetched.append((i, pos, end))
    if not fetched:
        return
    lo = min(p for _, p, _ in fetched)
    hi = max(e for _, _, e in fetched)
    for col in range(lo, hi):
        reads = [i for i, p, e in fetched if p <= col < e]
        if reads:
            yield col, reads


def bam_worker_reads(records, tid, start_pos, end_pos, vcf_handler, n_threads=1, stepper="samtools", max_depth=8000):
    """gretel/util.py:112-210 for every work block of util.py:294-301: a list (one entry per work block, in block
    order) of ``{read key: {"rank": int, "seq": [str, ...]}}`` dictionaries in first-seen order."""
    region = vcf_handler["region"]
    window_l = int(round((end_pos - start_pos) / float(n_threads)))                       # util.py:294
    blocks = []
    for window_i, window_pos in enumerate(range(start_pos, end_pos + 1, max(window_l, 1))):   # util.py:295-301
        blocks.append({"start": window_pos, "end": window_pos + window_l - 1, "i": window_i})
    out = []
    for work_block in blocks:
        reads = {}
        for col, in_col in pileup(records, tid, work_block["start"] - 1, work_block["end"], stepper, max_depth):   # :137
            if col + 1 > end_pos:                                                       # :139-141
                break
            if region[col + 1] != 1:                                                    # :143-145
                continue
            for ri in in_col:                                                           # :147
                rtid, pos, flag, name, cigar, seq = records[ri]
                one_or_two = 0                                                          # :149-158
                if flag & FPAIRED:
                    if flag & FREAD1:
                        one_or_two = 1
                    elif flag & FREAD2:
                        one_or_two = 2
                key = "%s_%s_%d" % (name, str(flag), one_or_two)                        # :160
                leftmost_1pos = pos + 1                                                 # :162
                if work_block["i"] == 0:                                                # :165-171
                    if leftmost_1pos < start_pos:
                        if pos + 1 + _query_alignment_length(cigar) < start_pos:
                            continue
                        leftmost_1pos = start_pos
                else:                                                                   # :172-176
                    if leftmost_1pos < work_block["start"]:
                        continue
                qpos, is_del = _at(records[ri], col)
                if is_del:                                                              # :180-183 ('-' x (|indel|+1))
                    sequence = "-"
                else:                                                                   # :184-190 (base + inserted bases)
                    sequence = seq[qpos] if qpos < len(seq) else "N"
                if key not in reads:                                                    # :196-203
                    reads[key] = {"rank": int(np.sum(region[1:leftmost_1pos])), "seq": []}
                reads[key]["seq"].append(sequence)                                      # :206
        out.append(reads)
    return out


def packed(blocks_reads):
    """The per-read support the pair expansion starts from (util.py:227-238): reads with at least two alleles,
    as a sorted list of ``(rank, tuple(codes))`` (the matrix does not depend on the order of the reads)."""
    out = []
    for reads in blocks_reads:
        for key, r in reads.items():
            if not len(r["seq"]) > 1:                                                   # :230
                continue
            support_seq = "".join(b[0] for b in r["seq"])                               # :238
            out.append((r["rank"], tuple(CODE.get(ch.upper(), CODE["N"]) for ch in support_seq)))
    return sorted(out)


def process_vcf_positions(positions, start_pos, end_pos):
    """gretel/util.py:388-414 over the 1-based POS column of the contig's VCF records (in file order)."""
    region = np.zeros(end_pos + 1, dtype=int)
    snp_reverse, snp_forward = {}, {}
    i = 0
    for pos in positions:
        if pos < start_pos or pos > end_pos:
            continue
        region[pos] = 1
        snp_reverse[i] = pos
        snp_forward[pos] = i
        i += 1
    return {"N": i, "snp_fwd": snp_forward, "snp_rev": snp_reverse, "region": region}
