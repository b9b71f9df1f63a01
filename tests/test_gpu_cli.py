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
