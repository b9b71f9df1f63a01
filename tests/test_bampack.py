"""The native multi-threaded BAM packer (hx_pack_bam, CPU code in libhanselx.so) must produce exactly
what the dependency-free Python packer produces, which in turn reproduces the reference's ingestion on its
own fixture (tests/test_oracle_golden.py)."""
import os

import numpy as np
import pytest

from gretel_b200 import bamio
from tests.bamwriter import write_bam


def _random_bam(path, rng, n_reads, ref_len=3000):
    refs = [("ctgA", ref_len), ("ctgB", ref_len)]
    reads = []
    for i in range(n_reads):
        tid = 0 if rng.random() < 0.9 else 1
        pos = int(rng.integers(0, ref_len - 50))
        cigar, qlen, rlen = [], 0, 0
        if rng.random() < 0.2:
            s = int(rng.integers(1, 8)); cigar.append(("S", s)); qlen += s
        for _ in range(int(rng.integers(1, 6))):
            m = int(rng.integers(5, 60)); cigar.append((rng.choice(["M", "=", "X"], p=[0.8, 0.1, 0.1]), m)); qlen += m; rlen += m
            r = rng.random()
            if r < 0.25:
                d = int(rng.integers(1, 12)); cigar.append(("D", d)); rlen += d
            elif r < 0.45:
                ins = int(rng.integers(1, 6)); cigar.append(("I", ins)); qlen += ins
            elif r < 0.5:
                n = int(rng.integers(5, 40)); cigar.append(("N", n)); rlen += n
        if cigar[-1][0] in "DIN":
            cigar.append(("M", 7)); qlen += 7; rlen += 7
        if rng.random() < 0.2:
            s = int(rng.integers(1, 8)); cigar.append(("S", s)); qlen += s
        if pos + rlen > ref_len:
            continue
        flag = int(rng.choice([0, 16, 99, 147, 65, 4, 256, 512, 1024, 2048], p=[.3, .3, .1, .1, .05, .03, .03, .03, .03, .03]))
        seq = "".join(rng.choice(list("ACGTN"), p=[.24, .24, .24, .24, .04], size=qlen))
        reads.append((tid, pos, flag, "r%d" % i, cigar, seq))
    reads.sort(key=lambda r: (r[0], r[1]))
    write_bam(path, refs, reads)
    return refs


@pytest.mark.parametrize("seed", range(4))
def test_native_packer_matches_python(tmp_path, seed):
    rng = np.random.default_rng(40 + seed)
    path = str(tmp_path / "t.bam")
    _random_bam(path, rng, 3000)
    snps = np.sort(rng.choice(np.arange(1, 3001), size=int(rng.integers(50, 400)), replace=False))
    for (start, end) in ((1, 3000), (500, 2200)):
        sel = [int(p) for p in snps if start <= p <= end]
        vh = {"N": len(sel), "snp_rev": {i: p for i, p in enumerate(sel)}}
        for stepper in ("samtools", "all", "nofilter"):
            exp = bamio.pack_bam(path, "ctgA", start, end, vh, stepper=stepper)
            for threads in (1, 3):
                got = bamio.pack_bam_native(path, "ctgA", start, end, vh, stepper=stepper, n_threads=threads)
                for a, b in zip(exp, got):
                    assert np.array_equal(a, b), (start, end, stepper, threads)
            assert len(exp[0]) > 50


def test_native_packer_on_reference_fixture(golden_dir):
    v = bamio.process_vcf(os.path.join(golden_dir, "ref_test.vcf.gz"), "hoot", 1, 20)
    exp = bamio.pack_bam(os.path.join(golden_dir, "ref_test.bam"), "hoot", 1, 20, v)
    got = bamio.pack_bam_native(os.path.join(golden_dir, "ref_test.bam"), "hoot", 1, 20, v, n_threads=2)
    for a, b in zip(exp, got):
        assert np.array_equal(a, b)
    assert list(got[0]) == [0, 0, 0, 0, 2]


def test_native_packer_errors(tmp_path, golden_dir):
    from gretel_b200 import _lib
    v = {"N": 1, "snp_rev": {0: 5}}
    with pytest.raises(_lib.HanselxError):
        bamio.pack_bam_native(str(tmp_path / "missing.bam"), "hoot", 1, 20, v)
    with pytest.raises(_lib.HanselxError):
        bamio.pack_bam_native(os.path.join(golden_dir, "ref_test.bam"), "nope", 1, 20, v)
    (tmp_path / "junk.bam").write_bytes(b"not a bam file at all")
    with pytest.raises(_lib.HanselxError):
        bamio.pack_bam_native(str(tmp_path / "junk.bam"), "hoot", 1, 20, v)


def _py_coverage(path, contig, start0, end0):
    refs, recs = bamio.read_bam(path)
    tid = [n for n, _ in refs].index(contig)
    out = np.zeros((4, end0 - start0), dtype=np.uint32)
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    for r in recs:
        if r.tid != tid or r.pos < 0:
            continue
        rpos, qpos = r.pos, 0
        for op, ln in r.cigar:
            if op in (0, 7, 8):
                for j in range(ln):
                    p = rpos + j
                    if start0 <= p < end0 and r.seq[qpos + j] in code:
                        out[code[r.seq[qpos + j]], p - start0] += 1
                rpos += ln; qpos += ln
            elif op in (2, 3):
                rpos += ln
            elif op in (1, 4):
                qpos += ln
    return out


def test_snpper_matches_python(tmp_path, golden_dir):
    """gretel/snpper.py:30-41 on the native reader: per-position base counts and the called sites."""
    import io
    from gretel_b200 import snpper
    rng = np.random.default_rng(9)
    path = str(tmp_path / "s.bam")
    _random_bam(path, rng, 2500)
    for (s0, e0) in ((0, 3000), (400, 1900)):
        exp = _py_coverage(path, "ctgA", s0, e0)
        for th in (1, 3):
            assert np.array_equal(snpper.count_coverage(path, "ctgA", s0, e0, n_threads=th, device=None), exp)
    assert snpper.contig_length(path, "ctgB") == 3000
    buf = io.StringIO()
    assert snpper.main(["--bam", path, "--contig", "ctgA", "-s", "401", "-e", "1900", "--depth", "2", "--cpu"], out=buf) == 0
    lines = buf.getvalue().strip().split("\n")
    assert lines[0] == "##fileformat=VCFv4.2"
    exp = _py_coverage(path, "ctgA", 400, 1900)
    sites = [i + 401 for i in np.nonzero((exp > 2).sum(axis=0) > 1)[0]]
    assert [int(l.split("\t")[1]) for l in lines[1:]] == sites and len(sites) > 10
    assert lines[1].split("\t")[3:] == ["A", "C,T,G", "0", ".", "INFO"]
    # the reference's toy BAM: reads 1-4 disagree at positions 1 and 2 of 'hoot' (A, C, T, T)
    buf = io.StringIO()
    snpper.main(["--bam", os.path.join(golden_dir, "ref_test.bam"), "--contig", "hoot", "--cpu"], out=buf)
    assert [int(l.split("\t")[1]) for l in buf.getvalue().strip().split("\n")[1:]] == [1, 2, 10]


@pytest.mark.gpu
def test_snpper_gpu_histogram(tmp_path, golden_dir):
    """gretel-snpper with the per-position histogram on the GPU (hx_count_coverage_gpu, coverage.cu): equal to the
    CPU counts on random BAMs (indels, clips, skips, several threads, sub-regions), on a deep synthetic BAM, and the
    reference's toy BAM gives the reference's sites (gretel/snpper.py:29-41)."""
    import io
    from gretel_b200 import snpper, synth
    rng = np.random.default_rng(19)
    path = str(tmp_path / "g.bam")
    _random_bam(path, rng, 4000)
    for (s0, e0) in ((0, 3000), (400, 1900), (2999, 3000)):
        exp = _py_coverage(path, "ctgA", s0, e0)
        for th in (1, 4):
            assert np.array_equal(snpper.count_coverage(path, "ctgA", s0, e0, n_threads=th, device=0), exp)
    w = synth.scaled(synth.WORKLOADS["metagenome"], 200_000)
    d = synth.generate(w)
    deep = str(tmp_path / "deep.bam")
    synth.write_bam(deep, d, w)
    cpu = snpper.count_coverage(deep, "ctg", 0, w.genome_len, n_threads=4, device=None)
    gpu = snpper.count_coverage(deep, "ctg", 0, w.genome_len, n_threads=4, device=0)
    assert np.array_equal(cpu, gpu) and int(gpu.sum()) > 25_000_000
    buf = io.StringIO()
    snpper.main(["--bam", os.path.join(golden_dir, "ref_test.bam"), "--contig", "hoot"], out=buf)
    assert [int(l.split("\t")[1]) for l in buf.getvalue().strip().split("\n")[1:]] == [1, 2, 10]


@pytest.mark.parametrize("layout", ["aligned", "cut_anywhere", "tiny_blocks"])
def test_record_scan_over_many_blocks(tmp_path, layout):
    """The packer finds the records of a wave with a speculative parallel walk from BGZF block starts (htslib ends a
    block rather than cut a record) and falls back to the serial chain where the real chain does not arrive at a block
    start.  Files of hundreds of blocks in both layouts, 1 and 8 threads: identical packed reads, equal to the
    Python packer's on a sample window."""
    rng = np.random.default_rng(7)
    G, n = 60_000, 40_000
    starts = np.sort(rng.integers(0, G - 150, size=n))
    bases = np.array(list("ACGT"))
    seqs = bases[rng.integers(0, 4, size=(n, 150))]
    reads = [(0, int(s0), 0, "r%d" % i, [("M", 100), ("D", 2), ("M", 50)] if i % 7 == 0 else [("M", 150)], "".join(seqs[i]))
             for i, s0 in enumerate(starts)]
    path = str(tmp_path / "big.bam")
    write_bam(path, [("ctg", G)], reads, block={"aligned": 30000, "cut_anywhere": 30011, "tiny_blocks": 700}[layout],
              align=layout != "cut_anywhere")
    snps = np.sort(rng.choice(np.arange(1, G + 1), size=4000, replace=False))
    vh = {"N": len(snps), "snp_rev": {i: int(p) for i, p in enumerate(snps)}}
    one = bamio.pack_bam_native(path, "ctg", 1, G, vh, n_threads=1)
    many = bamio.pack_bam_native(path, "ctg", 1, G, vh, n_threads=8)
    for a, b in zip(one, many):
        assert np.array_equal(a, b)
    assert len(one[0]) > 30_000
    lo, hi = 20_000, 22_000                                  # the slow Python packer on a window of the same file
    sel = [int(p) for p in snps if lo <= p <= hi]
    vw = {"N": len(sel), "snp_rev": {i: p for i, p in enumerate(sel)}}
    exp = bamio.pack_bam(path, "ctg", lo, hi, vw)
    got = bamio.pack_bam_native(path, "ctg", lo, hi, vw, n_threads=8)
    for a, b in zip(exp, got):
        assert np.array_equal(a, b)
