"""Config 1 of BASELINE.json: the reference's CLI flow on its own toy BAM + VCF, through the GPU
path, compared with the same flow driven by the oracle (outputs formatted as gretel/cmd.py:181-240)."""
import os

import numpy as np
import pytest

from gretel_b200 import bamio
from oracle import hansel_oracle as o

pytestmark = pytest.mark.gpu


def test_cli_outputs_match_oracle(tmp_path, golden_dir, capsys):
    from gretel_b200 import cmd
    bam, vcf = os.path.join(golden_dir, "ref_test.bam"), os.path.join(golden_dir, "ref_test.vcf.gz")
    rc = cmd.main([bam, vcf, "hoot", "-s", "1", "-e", "20", "-p", "6", "-o", str(tmp_path),
                   "--dumpmatrix", str(tmp_path / "m.npz"), "--dumpsnps", str(tmp_path / "snps.tsv")])
    assert rc == 0
    table = capsys.readouterr().out.strip().split("\n")
    assert table[0].split("\t") == ["i", "pos", "gap", "A", "C", "G", "T", "N", "-", "_", "tot"]
    assert table[2].split("\t") == ["1", "1", "1", "1", "1", "0", "2", "0", "0", "0", "4"]     # site 1: A C T T
    # oracle-driven expectation
    v = bamio.process_vcf(vcf, "hoot", 1, 20)
    rank, off, codes = bamio.pack_bam(bam, "hoot", 1, 20, v)
    ho = o.load_from_packed(rank, off, codes, v["N"])
    its, PATHS = o.recover(ho, v["N"], max_paths=6)
    key = list(PATHS)[0]
    fasta = open(tmp_path / "out.fasta").read().split("\n")
    assert fasta[0] == ">0__%.2f" % PATHS[key]["hp_current"][0]
    seq = ["N"] * 20
    for j, a in enumerate(PATHS[key]["hansel_path"][1:]):
        seq[v["snp_rev"][j] - 1] = a
    assert fasta[1] == "".join(seq)
    snp = open(tmp_path / "snp.fasta").read().split("\n")
    assert snp[1] == key[1:]
    crumbs = open(tmp_path / "gretel.crumbs").read().strip().split("\n")
    assert crumbs[0] == "# %d\t%d\t%d\t%.2f" % (v["N"], ho.n_crumbs, ho.n_slices, ho.L)
    p = PATHS[key]
    assert crumbs[1] == "%d\t%d\t%s\t%s\t%.2f" % (p["i_0"], p["n"], ",".join("%.2f" % x for x in p["hp_current"]),
                                                ",".join("%.2f" % x for x in p["hp_original"]), p["magnitude"])
    assert open(tmp_path / "snps.tsv").read().split("\n")[0] == "1\t1\t1"
    # --dumpmatrix round trip
    from gretel_b200.hansel import Hansel
    h2 = Hansel.load_hansel_dump(str(tmp_path / "m.npz"))
    assert h2.get_observation('T', 'T', 1, 2) == 2 and h2.n_crumbs == 9 and h2.L == 3


def test_cli_reports_gap(tmp_path, golden_dir):
    """cmd.py:85-118: a SNP that no read bridges aborts with status 1 (hoot:1-19 has no read over ... site 3 alone)."""
    from gretel_b200 import cmd
    bam, vcf = os.path.join(golden_dir, "ref_test.bam"), os.path.join(golden_dir, "ref_test.vcf.gz")
    # restricting to 1..19 leaves SNPs {1,2,10}; read5 then covers a single SNP and site 3 has no right neighbour
    v = bamio.process_vcf(vcf, "hoot", 1, 19)
    rank, off, codes = bamio.pack_bam(bam, "hoot", 1, 19, v)
    ho = o.load_from_packed(rank, off, codes, v["N"])
    exp_gaps = o.gap_check(ho, v["N"])
    rc = cmd.main([bam, vcf, "hoot", "-s", "1", "-e", "19", "-p", "2", "-o", str(tmp_path), "--quiet"])
    assert (rc == 1) == bool(exp_gaps)


def test_pipeline_on_synthetic_bam(tmp_path):
    """BAM file -> native packer -> GPU ingestion -> resident recovery -> crumbs, against the same flow
    through the Python packer and the oracle (3 strains over 40 SNPs, 1500 reads with deletions)."""
    import gzip
    from gretel_b200 import cmd
    from tests.bamwriter import write_bam
    rng = np.random.default_rng(123)
    G, N, R, L = 1200, 40, 1500, 100
    sites = 20 + np.cumsum(rng.integers(5, 50, size=N))       # neighbours always bridged by a 100 bp read
    assert sites[-1] < G - 20
    ref = rng.choice(list("ACGT"), size=G)
    strains = []
    for s in range(3):
        g = ref.copy()
        for p in sites:
            if rng.random() < 0.6:
                g[p] = rng.choice([b for b in "ACGT" if b != ref[p]])
        strains.append(g)
    reads = []
    for i in range(R):
        st = strains[rng.choice(3, p=[0.5, 0.3, 0.2])]
        pos = int(rng.integers(0, G - L))
        if rng.random() < 0.15:                               # a 3-base deletion in the middle
            cigar = [("M", 40), ("D", 3), ("M", L - 43)]
            seq = "".join(st[pos:pos + 40]) + "".join(st[pos + 43:pos + L])
        else:
            cigar = [("M", L)]
            seq = "".join(st[pos:pos + L])
        reads.append((0, pos, 0, "r%d" % i, cigar, seq))
    reads.sort(key=lambda r: r[1])
    bam, vcf = str(tmp_path / "s.bam"), str(tmp_path / "s.vcf.gz")
    write_bam(bam, [("ctg", G)], reads, block=20000)
    with gzip.open(vcf, "wt") as fh:
        fh.write("##fileformat=VCFv4.2\n")
        for p in sites:
            fh.write("ctg\t%d\t.\tA\tC,T,G\t0\t.\tINFO\n" % (p + 1))
    out = tmp_path / "out"
    out.mkdir()
    assert cmd.main([bam, vcf, "ctg", "-s", "1", "-e", str(G), "-p", "4", "-o", str(out), "--quiet", "-@", "2"]) == 0
    # oracle flow
    v = bamio.process_vcf(vcf, "ctg", 1, G)
    assert v["N"] == N
    rank, off, codes = bamio.pack_bam(bam, "ctg", 1, G, v)
    ho = o.load_from_packed(rank, off, codes, N)
    assert o.gap_check(ho, N) == []
    its, PATHS = o.recover(ho, N, max_paths=4)
    crumbs = open(out / "gretel.crumbs").read().strip().split("\n")
    assert crumbs[0] == "# %d\t%d\t%d\t%.2f" % (N, ho.n_crumbs, ho.n_slices, ho.L)
    exp_rows = []
    for key in sorted(PATHS, key=lambda x: PATHS[x]["hp_current"][0], reverse=True):
        p = PATHS[key]
        exp_rows.append("%d\t%d\t%s\t%s\t%.2f" % (p["i_0"], p["n"], ",".join("%.2f" % x for x in p["hp_current"]),
                                                  ",".join("%.2f" % x for x in p["hp_original"]), p["magnitude"]))
    assert crumbs[1:] == exp_rows
    snp = open(out / "snp.fasta").read().strip().split("\n")[1::2]
    assert snp == [k[1:] for k in sorted(PATHS, key=lambda x: PATHS[x]["i_0"])]
    # the dominant strain is recovered first
    assert snp[0] == "".join(strains[0][sites])
