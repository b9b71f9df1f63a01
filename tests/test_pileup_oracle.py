"""The read packers (the native streaming packer hx_pack_bam_ex and the dependency-free Python packer) against a
literal column-by-column restatement of the reference's pileup loop (oracle/pileup_oracle.py, gretel/util.py:112-210):
hand-written SAM cases for every rule the reference applies, plus random BAMs.  The CIGAR-walk packers never see
a pileup column, so agreeing with the column walk on these cases is the parity evidence for SURVEY row I2."""
import os

import numpy as np
import pytest

from gretel_b200 import bamio
from oracle import pileup_oracle as po
from tests.bamwriter import write_bam
from tests.test_bampack import _random_bam

_OPS = {c: i for i, c in enumerate("MIDNSHP=X")}


def _parse_cigar(s):
    out, n = [], ""
    for ch in s:
        if ch.isdigit():
            n += ch
        else:
            out.append((ch, int(n)))
            n = ""
    return out


def _sam(text):
    """'name flag contig pos1 cigar seq' lines -> reads for tests.bamwriter (contigs: ctg = 0, other = 1)."""
    reads = []
    for line in text.strip().splitlines():
        name, flag, ctg, pos1, cigar, seq = line.split()
        reads.append(({"ctg": 0, "other": 1}[ctg], int(pos1) - 1, int(flag), name, _parse_cigar(cigar), seq.upper()))   # BAM holds upper case
    return reads


def _oracle_records(reads):
    return [(tid, pos, flag, name, [(_OPS[op], ln) for op, ln in cigar], seq) for tid, pos, flag, name, cigar, seq in reads]


def _as_sorted(packed):
    rank, off, codes = packed
    return sorted((int(rank[i]), tuple(int(c) for c in codes[off[i]:off[i + 1]])) for i in range(len(rank)))


def _check(tmp_path, reads, positions, start, end, steppers=("samtools", "all"), max_depth=8000, n_threads=(1, 2),
           ref_len=200):
    path = str(tmp_path / "case.bam")
    write_bam(path, [("ctg", ref_len), ("other", ref_len)], reads)
    vh = po.process_vcf_positions(positions, start, end)
    recs = _oracle_records(reads)
    results = []
    for stepper in steppers:
        want = po.packed(po.bam_worker_reads(recs, 0, start, end, vh, n_threads=1, stepper=stepper, max_depth=max_depth))
        for nt in n_threads:
            # the reference's own work-block split must not change what is extracted (util.py:162-176)
            if nt > 1:
                split = po.packed(po.bam_worker_reads(recs, 0, start, end, vh, n_threads=nt, stepper=stepper, max_depth=0))
                whole = po.packed(po.bam_worker_reads(recs, 0, start, end, vh, n_threads=1, stepper=stepper, max_depth=0))
                assert split == whole, (stepper, nt)
        got_py = _as_sorted(bamio.pack_bam(path, "ctg", start, end, vh, stepper=stepper, max_depth=max_depth))
        assert got_py == want, ("python packer", stepper)
        for threads in (1, 4):
            got = _as_sorted(bamio.pack_bam_native(path, "ctg", start, end, vh, stepper=stepper, n_threads=threads,
                                                   max_depth=max_depth))
            assert got == want, ("native packer", stepper, threads)
        results.append(want)
    return results


def test_reference_fixture_reads(golden_dir):
    """The reference's own fixture through the column walk: the per-read support of SURVEY section 4's hand trace."""
    refs, recs = bamio.read_bam(os.path.join(golden_dir, "ref_test.bam"))
    records = [(r.tid, r.pos, r.flag, r.name, list(r.cigar), r.seq) for r in recs]
    vh = bamio.process_vcf(os.path.join(golden_dir, "ref_test.vcf.gz"), "hoot", 1, 20)
    for nt in (1, 2):
        got = po.packed(po.bam_worker_reads(records, 0, 1, 20, vh, n_threads=nt))
        assert got == sorted([(0, (0, 0, 0)), (0, (1, 1, 1)), (0, (3, 3)), (0, (3, 3)), (2, (2, 2))])


def test_indels_skips_and_clips(tmp_path):
    # SNPs at 5, 10, 12, 20, 30, 41
    reads = _sam("""
        plain     0  ctg 3  40M        ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT
        del       0  ctg 3  6M4D30M    ACGTACGTACGTACGTACGTACGTACGTACGTACGT
        ins       0  ctg 3  7M3I30M    ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT
        skip      0  ctg 3  5M20N15M   ACGTACGTACGTACGTACGT
        clipped   0  ctg 8  4S10M3S    TTTTACGTACGTACGGG
        hardclip  0  ctg 8  5H25M      ACGTACGTACGTACGTACGTACGTA
        eqx       0  ctg 9  3=2X20=    ACGTACGTACGTACGTACGTACGTA
        endsdel   0  ctg 9  2M2D20M    ACGTACGTACGTACGTACGTAC
        lower     0  ctg 4  30M        acgtacgtacgtacgtacgtacgtacgtac
        ambig     0  ctg 4  30M        ACGTACRYACGTACGTNCGTACGTACGTAC
        other     0  other 3 40M       ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT
    """)
    _check(tmp_path, reads, [5, 10, 12, 20, 30, 41], 1, 60)


def test_reads_before_start_and_past_end(tmp_path):
    reads = _sam("""
        early_in   0 ctg 2   30M      ACGTACGTACGTACGTACGTACGTACGTAC
        early_out  0 ctg 2   8M       ACGTACGT
        early_del  0 ctg 2   4M30D10M ACGTACGTACGTAC
        inside     0 ctg 15  20M      ACGTACGTACGTACGTACGT
        crosses    0 ctg 30  30M      ACGTACGTACGTACGTACGTACGTACGTAC
        beyond     0 ctg 45  10M      ACGTACGTAC
    """)
    # early_del: reference_start + 1 + query_alignment_length < start_pos although its deletion reaches the
    # region (util.py:168 measures the query, not the reference span) -> dropped
    _check(tmp_path, reads, [8, 12, 16, 20, 25, 33, 38, 44, 47, 52], 11, 44)
    _check(tmp_path, reads, [8, 12, 16, 20, 25, 33, 38, 44, 47, 52], 1, 60)


def test_steppers_and_mates(tmp_path):
    seq = "ACGTACGTACGTACGTACGTACGTACGTAC"
    reads = _sam("""
        pair   99   ctg 3  30M %s
        pair   147  ctg 5  30M %s
        orphan 65   ctg 4  30M %s
        orphan 129  ctg 6  30M %s
        unmap  4    ctg 7  30M %s
        second 256  ctg 7  30M %s
        qcfail 512  ctg 8  30M %s
        dup    1024 ctg 8  30M %s
        suppl  2048 ctg 9  30M %s
        single 16   ctg 9  30M %s
    """ % ((seq,) * 10))
    got = _check(tmp_path, reads, [5, 10, 12, 20, 30], 1, 60, steppers=("samtools", "all", "nofilter"))
    assert len(got[0]) < len(got[1]) < len(got[2])          # orphans only with "all", flagged reads only unfiltered


def test_depth_cap(tmp_path):
    """pysam's max_depth: with a cap of 3, the fourth read of a start position is dropped while three reads are
    buffered, the first read of a new start position always gets in, and leaving reads free their place."""
    seq = "ACGTACGTACGTACGTACGT"
    lines = []
    for i in range(6):
        lines.append("a%d 0 ctg 3 20M %s" % (i, seq))
    lines.append("b0 0 ctg 5 20M %s" % seq)
    lines.append("b1 0 ctg 5 20M %s" % seq)
    for i in range(3):
        lines.append("c%d 0 ctg 40 20M %s" % (i, seq))
    reads = _sam("\n".join(lines))
    pos = [4, 8, 12, 16, 21, 42, 48, 55]
    capped = _check(tmp_path, reads, pos, 1, 80, steppers=("all",), max_depth=3)[0]
    full = _check(tmp_path, reads, pos, 1, 80, steppers=("all",), max_depth=0)[0]
    assert len(full) == 11 and len(capped) == 3 + 1 + 3      # a0-a2, b0 (first of its position), c0-c2


def test_duplicate_vcf_positions(tmp_path):
    """A VCF that repeats a POS: the reference counts the record in N but marks one region position (util.py:402-406),
    so ranks and alleles follow the unique positions."""
    reads = _sam("""
        r1 0 ctg 2  30M ACGTACGTACGTACGTACGTACGTACGTAC
        r2 0 ctg 9  30M ACGTACGTACGTACGTACGTACGTACGTAC
    """)
    vh = po.process_vcf_positions([5, 10, 10, 20, 20, 30], 1, 60)
    assert vh["N"] == 6 and int(vh["region"].sum()) == 4
    assert bamio.snp_positions(vh) == [5, 10, 20, 30]
    _check(tmp_path, reads, [5, 10, 10, 20, 20, 30], 1, 60)


def test_records_sharing_a_key_are_kept_apart(tmp_path):
    """Deliberate deviation: the reference keys reads by name + flag + mate (util.py:160), so two records with the
    same key are glued into one support sequence (alleles interleaved column by column - not a haplotype).  The
    packers keep every BAM record as its own read."""
    reads = _sam("""
        twin 0 ctg 3  20M ACGTACGTACGTACGTACGT
        twin 0 ctg 7  20M ACGTACGTACGTACGTACGT
    """)
    path = str(tmp_path / "twin.bam")
    write_bam(path, [("ctg", 100), ("other", 100)], reads)
    vh = po.process_vcf_positions([5, 10, 12, 20, 24], 1, 60)
    merged = po.packed(po.bam_worker_reads(_oracle_records(reads), 0, 1, 60, vh))
    assert len(merged) == 1 and len(merged[0][1]) == 8        # 4 + 4 alleles under one key
    ours = _as_sorted(bamio.pack_bam_native(path, "ctg", 1, 60, vh))
    assert ours == sorted([(0, (0, 1, 3, 3)), (1, (3, 1, 1, 1))]) or len(ours) == 2
    assert ours == _as_sorted(bamio.pack_bam(path, "ctg", 1, 60, vh))


@pytest.mark.parametrize("seed", range(3))
def test_random_bams_against_the_column_walk(tmp_path, seed):
    rng = np.random.default_rng(70 + seed)
    path = str(tmp_path / "r.bam")
    _random_bam(path, rng, 400, ref_len=600)
    refs, recs = bamio.read_bam(path)
    records = [(r.tid, r.pos, r.flag, r.name, list(r.cigar), r.seq) for r in recs]
    positions = sorted(int(p) for p in rng.choice(np.arange(1, 601), size=90, replace=False))
    for start, end in ((1, 600), (120, 430)):
        vh = po.process_vcf_positions(positions, start, end)
        for stepper in ("samtools", "all"):
            for depth in (0, 25):
                want = po.packed(po.bam_worker_reads(records, 0, start, end, vh, stepper=stepper, max_depth=depth))
                assert len(want) > 20
                assert _as_sorted(bamio.pack_bam(path, "ctgA", start, end, vh, stepper=stepper, max_depth=depth)) == want
                assert _as_sorted(bamio.pack_bam_native(path, "ctgA", start, end, vh, stepper=stepper, n_threads=3,
                                                        max_depth=depth)) == want


def test_long_cigar_in_cg_tag(tmp_path):
    """Alignments with more than 65535 CIGAR operations keep the real CIGAR in the CG:B,I tag behind the placeholder
    <l_seq>S<ref_len>N (SAM spec 4.2.2)."""
    import struct
    from tests import bamwriter
    seq = "ACGTACGTACGTACGTACGTACGTACGTAC"
    real = [("M", 10), ("D", 2), ("M", 20)]
    path = str(tmp_path / "cg.bam")
    cg = b"CGBI" + struct.pack("<I", len(real)) + b"".join(struct.pack("<I", (ln << 4) | _OPS[op]) for op, ln in real)
    bamwriter.write_bam(path, [("ctg", 200)], [(0, 4, 0, "long", [("S", len(seq)), ("N", 32)], seq)], aux=[cg])
    vh = po.process_vcf_positions([6, 12, 15, 16, 30], 1, 100)
    want = po.packed(po.bam_worker_reads([(0, 4, 0, "long", [(_OPS[o], l) for o, l in real], seq)], 0, 1, 100, vh))
    assert _as_sorted(bamio.pack_bam_native(path, "ctg", 1, 100, vh)) == want and len(want) == 1
