"""Command-line driver with the reference's flags and outputs (gretel/cmd.py:11-240), on the B200 path.

    python -m gretel_b200 <bam> <vcf.gz> <contig> [-s START] [-e END] [-p PATHS] [-o OUT] ...

Everything that computes (ingestion, gap check counts, recovery, reweighting) runs in the CUDA
library; this file is host-side bookkeeping and output formatting only: PATHS de-duplication
(cmd.py:164-179), out.fasta / snp.fasta (cmd.py:181-221) and gretel.crumbs (cmd.py:223-240,
docs/protocol.rst:48-77).
"""
from __future__ import annotations

import argparse
import sys

from . import gretel, util

__version__ = "0.0.94+b200"


def read_first_fasta_record(path):
    """First sequence of a FASTA file (the reference uses pysam.FastaFile, util.py:337-351)."""
    seq = []
    seen = False
    with open(path) as fh:
        for line in fh:
            if line.startswith(">"):
                if seen:
                    break
                seen = True
                continue
            if seen:
                seq.append(line.strip())
    return "".join(seq)


def build_parser():
    p = argparse.ArgumentParser(prog="gretel", description="Gretel: A metagenomic haplotyper (B200 hot path).")
    p.add_argument("bam")
    p.add_argument("vcf")
    p.add_argument("contig")
    p.add_argument("-s", "--start", type=int, default=1)
    p.add_argument("-e", "--end", type=int, default=-1)
    p.add_argument("-p", "--paths", type=int, default=100)
    p.add_argument("--master", default=None)
    p.add_argument("--gapchar", default="N")
    p.add_argument("--delchar", default="")
    p.add_argument("--quiet", default=False, action="store_true")
    p.add_argument("-o", "--out", default=".")
    p.add_argument("-@", "--threads", type=int, default=1)
    p.add_argument("--dumpmatrix", type=str, default=None)
    p.add_argument("--dumpsnps", type=str, default=None)
    p.add_argument("--pepper", action="store_true")
    p.add_argument("--device", type=int, default=None)
    p.add_argument("--version", action="version", version="%(prog)s " + __version__)
    return p


def snp_table(hansel, vcf_h, out=sys.stdout):
    """cmd.py:123-145."""
    out.write("i\tpos\tgap\tA\tC\tG\tT\tN\t-\t_\ttot\n")
    last_rev = 0
    for i in range(0, vcf_h["N"] + 1):
        m = {str(k): v for k, v in hansel.get_counts_at(i).items()}
        snp_rev = vcf_h["snp_rev"][i - 1] if i > 0 else 0
        out.write("%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n" % (
            i, snp_rev, snp_rev - last_rev, m.get("A", 0), m.get("C", 0), m.get("G", 0), m.get("T", 0),
            m.get("N", 0), m.get("-", 0), m.get("_", 0), m.get("total", 0)))
        last_rev = snp_rev


def write_outputs(dirn, PATHS, hansel, vcf_h, start, end, master=None, gapchar="N", delchar=""):
    """cmd.py:181-240: out.fasta, snp.fasta, gretel.crumbs."""
    dirn = dirn.rstrip("/") + "/"
    master_seq = read_first_fasta_record(master) if master else [' '] * end
    with open(dirn + "out.fasta", "w") as fasta, open(dirn + "snp.fasta", "w") as hfasta:
        for key in sorted(PATHS, key=lambda x: PATHS[x]["i_0"]):
            p = PATHS[key]
            path, i = p["hansel_path"], p["i_0"]
            seq = list(master_seq[:])
            for j, mallele in enumerate(path[1:]):
                pos = vcf_h["snp_rev"][j]
                seq[pos - 1] = delchar if mallele == hansel.symbols_d["-"] else mallele
            to_write = "".join(str(x) for x in seq[start - 1:end])
            if not master:
                to_write = to_write.replace(' ', gapchar)
            fasta.write(">%d__%.2f\n%s\n" % (i, p["hp_current"][0], to_write))
            hfasta.write(">%d__%.2f\n%s\n" % (i, p["hp_current"][0], "".join(str(x) for x in path[1:])))
    with open(dirn + "gretel.crumbs", "w") as crumbs:
        crumbs.write("# %d\t%d\t%d\t%.2f\n" % (vcf_h["N"], hansel.n_crumbs, hansel.n_slices, hansel.L))
        for key in sorted(PATHS, key=lambda x: PATHS[x]["hp_current"][0], reverse=True):
            p = PATHS[key]
            crumbs.write("%d\t%d\t%s\t%s\t%.2f\n" % (
                p["i_0"], p["n"], ",".join("%.2f" % x for x in p["hp_current"]),
                ",".join("%.2f" % x for x in p["hp_original"]), p["magnitude"]))


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.end == -1:
        args.end = util.get_ref_len_from_bam(args.bam, args.contig)
        sys.stderr.write("[NOTE] Setting end_pos to %d" % args.end)
    vcf_h = util.process_vcf(args.vcf, args.contig, args.start, args.end)
    if args.dumpsnps:
        with open(args.dumpsnps, "w") as fh:
            for k in sorted(vcf_h["snp_fwd"].keys()):
                fh.write("%d\t%d\t%d\n" % (vcf_h["snp_fwd"][k] + 1, k, k - args.start + 1))
    hansel = util.load_from_bam(args.bam, args.contig, args.start, args.end, vcf_h, n_threads=args.threads,
                                stepper="all" if args.pepper else "samtools", device=args.device)
    if args.dumpmatrix:
        hansel.save_hansel_dump(args.dumpmatrix)
    gaps = gretel.gap_check(hansel, vcf_h["N"])
    if gaps:
        i = gaps[0]
        sys.stderr.write("[FAIL] Unable to recover pairwise evidence concerning SNP #%d at position %d\n" % (
            i, vcf_h["snp_rev"][i - 1] if i > 0 else 0))
        return 1
    if not args.quiet:
        snp_table(hansel, vcf_h)
    _, PATHS = gretel.recover(hansel, vcf_h["N"], max_paths=args.paths, min_remove=0.01)
    write_outputs(args.out, PATHS, hansel, vcf_h, args.start, args.end, master=args.master,
                  gapchar=args.gapchar, delchar=args.delchar)
    return 0


if __name__ == "__main__":
    sys.exit(main())
