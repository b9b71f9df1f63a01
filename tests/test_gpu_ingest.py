"""Parity of the CUDA ingestion path (through the C ABI) against the oracle."""
import os

import numpy as np
import pytest

from gretel_b200 import synth

pytestmark = pytest.mark.gpu


def _gpu_band(rank, off, codes, N, W, kernel=0):
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
    h.set_ingest_kernel(kernel)
    totals = h.ingest_packed(rank, off, codes)
    band = h.band()
    h.close()
    return band, totals


def test_reference_golden_through_gpu(golden_dir):
    """tests/test_test.py:33-52 of the reference, through load_from_bam on the GPU."""
    from gretel_b200 import util
    for threads in (1, 2):
        v = util.process_vcf(os.path.join(golden_dir, "ref_test.vcf.gz"), "hoot", 1, 20)
        h = util.load_from_bam(os.path.join(golden_dir, "ref_test.bam"), "hoot", 1, 20, v, n_threads=threads)
        assert h.n_slices == 5
        assert h.n_crumbs == 9
        assert h.L > 0
        g = h.get_observation
        assert g('_', 'A', 0, 1) == 1
        assert g('A', 'A', 1, 2) == 1
        assert g('A', 'A', 1, 3) == 1
        assert g('A', 'A', 1, 4) == 0
        assert g('C', 'C', 1, 2) == 1
        assert g('C', 'C', 1, 3) == 1
        assert g('C', 'C', 1, 4) == 0
        assert g('T', 'T', 1, 2) == 2
        assert g('G', 'G', 1, 2) == 0
        assert g('G', 'G', 2, 3) == 0
        assert g('G', 'G', 3, 4) == 1
        assert g('G', '_', 4, 5) == 1
        assert h.to_dense().sum() == 14
        assert h.L == 3


@pytest.mark.parametrize("kernel", [1, 2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("seed", range(8))
def test_random_packed_bit_exact(c_oracle, seed, kernel):
    rng = np.random.default_rng(500 + seed)
    N = int(rng.integers(2, 200))
    R = int(rng.integers(1, 3000))
    mk = int(rng.integers(2, 40))
    rank, off, codes = synth.random_packed(rng, N, R, mk, p_special=0.2, sort=bool(seed % 2))
    W = max(1, min(mk, N) - 1) if seed % 3 else min(N + 1, 64)
    ref, rt = c_oracle.ingest(rank, off, codes, N, W)
    band, totals = _gpu_band(rank, off, codes, N, W, kernel)
    assert totals == tuple(int(x) for x in rt)
    assert np.array_equal(band, ref.astype(np.float32))


@pytest.mark.parametrize("kernel", [1, 2, 3, 4, 5, 6, 7])
def test_edge_cases(c_oracle, kernel):
    # empty input, reads with k<2, N=2 (start rule beats end rule), reads ending on the last SNP
    cases = []
    cases.append((np.zeros(0, np.int32), np.zeros(1, np.int64), np.zeros(0, np.uint8), 5, 3))
    cases.append((np.array([0, 1, 2], np.int32), np.array([0, 1, 1, 2], np.int64), np.array([0, 1], np.uint8), 5, 3))
    cases.append((np.array([0, 0], np.int32), np.array([0, 2, 4], np.int64), np.array([0, 1, 4, 2], np.uint8), 2, 1))
    cases.append((np.array([0, 2, 3], np.int32), np.array([0, 5, 8, 10], np.int64),
                  np.array([0, 1, 2, 3, 5, 6, 0, 4, 4, 1], np.uint8), 5, 4))
    for rank, off, codes, N, W in cases:
        ref, rt = c_oracle.ingest(rank, off, codes, N, W)
        band, totals = _gpu_band(rank, off, codes, N, W, kernel)
        assert totals == tuple(int(x) for x in rt)
        assert np.array_equal(band, ref.astype(np.float32))


def test_bad_reads_raise():
    from gretel_b200 import _lib
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, 5, band_w=2)
    with pytest.raises(_lib.HanselxError):
        h.ingest_packed(np.array([4], np.int32), np.array([0, 3], np.int64), np.array([0, 1, 2], np.uint8))
    h2 = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, 5, band_w=1)
    with pytest.raises(_lib.HanselxError):
        h2.ingest_packed(np.array([0], np.int32), np.array([0, 3], np.int64), np.array([0, 1, 2], np.uint8))
    h3 = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, 5, band_w=4)
    with pytest.raises(_lib.HanselxError):
        h3.ingest_packed(np.array([0], np.int32), np.array([0, 3], np.int64), np.array([0, 9, 2], np.uint8))


@pytest.mark.parametrize("name,n_reads", [("hiv", 200_000), ("metagenome", 300_000), ("ont", 600)])
@pytest.mark.parametrize("kernel", [0, 1, 3, 4, 5, 6, 7])
def test_workloads_bit_exact(c_oracle, name, n_reads, kernel):
    """Config 2 at full size, configs 3/4 at sizes the C oracle finishes in seconds."""
    w = synth.scaled(synth.WORKLOADS[name], n_reads)
    d = synth.generate(w)
    W = d["max_k"] - 1
    ref, rt = c_oracle.ingest(d["rank"], d["off"], d["codes"], w.n_snps, W)
    band, totals = _gpu_band(d["rank"], d["off"], d["codes"], w.n_snps, W, kernel)
    assert totals == tuple(int(x) for x in rt)
    assert np.array_equal(band, ref.astype(np.float32))


def test_scalar_surface_and_spill():
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, 6, band_w=2)
    h.add_observation('A', 'C', 1, 2)
    h.add_observation('A', 'C', 1, 2)
    h.add_observation('A', 'C', 1, 5)           # outside the band -> host spill
    h.add_observation('G', 'G', 3, 3)           # diagonal -> spill
    assert h.get_observation('A', 'C', 1, 2) == 2
    assert h.get_observation('A', 'C', 1, 5) == 1
    assert h.get_observation('G', 'G', 3, 3) == 1
    assert h.get_observation('T', 'T', 2, 3) == 0
    assert h.reweight_observation('A', 'C', 1, 2, 0.25) == 0.5
    assert h.get_observation('A', 'C', 1, 2) == 1.5
    c = h.copy()
    c.add_observation('A', 'C', 1, 2)
    assert h.get_observation('A', 'C', 1, 2) == 1.5 and c.get_observation('A', 'C', 1, 2) == 2.5
    d = h.to_dense()
    assert d.shape == (7, 7, 8, 8) and d[0, 1, 1, 2] == 1.5 and d[0, 1, 1, 5] == 1


def test_linearity_and_order_independence():
    """Size-independent properties at a size the oracle is not asked to match:
    ingest(A)+ingest(B) == ingest(A||B) == ingest(shuffled A||B); sum == crumbs+sentinels."""
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    w = synth.scaled(synth.WORKLOADS["metagenome"], 2_000_000)
    d = synth.generate(w)
    W = d["max_k"] - 1
    R = len(d["rank"])
    whole, t_whole = _gpu_band(d["rank"], d["off"], d["codes"], w.n_snps, W)
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, w.n_snps, band_w=W)
    half = R // 2
    off = d["off"]
    h.ingest_packed(d["rank"][:half], off[:half + 1], d["codes"])
    t2 = h.ingest_packed(d["rank"][half:], off[half:], d["codes"])
    assert t2 == t_whole
    assert np.array_equal(h.band(), whole)
    assert float(whole.astype(np.float64).sum()) == t_whole[1] + t_whole[3]
    # generic kernel on a permutation of the reads
    rng = np.random.default_rng(3)
    perm = rng.permutation(R)
    k = np.diff(off)
    poff = np.zeros(R + 1, np.int64)
    np.cumsum(k[perm], out=poff[1:])
    idx = np.repeat(off[:-1][perm], k[perm]) + (np.arange(poff[-1]) - np.repeat(poff[:-1], k[perm]))
    pb, pt = _gpu_band(d["rank"][perm], poff, d["codes"][idx], w.n_snps, W)
    assert pt == t_whole and np.array_equal(pb, whole)


@pytest.mark.parametrize("name,n_reads", [("hiv", 30_000), ("metagenome", 100_000), ("ont", 300)])
def test_compact_wire_format(c_oracle, name, n_reads):
    """hx_ingest_host_compact (uint16 SNP counts + nibble codes, offsets rebuilt by a device scan)."""
    from gretel_b200 import util
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    w = synth.scaled(synth.WORKLOADS[name], n_reads)
    d = synth.generate(w)
    W = d["max_k"] - 1
    klen, codes4, n_codes = util.compact_packed(d["off"], d["codes"])
    assert n_codes == len(d["codes"]) and len(codes4) == (n_codes + 1) // 2
    ref, rt = c_oracle.ingest(d["rank"], d["off"], d["codes"], w.n_snps, W)
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, w.n_snps, band_w=W)
    totals = h.ingest_packed_compact(d["rank"], klen, codes4, n_codes)
    assert totals == tuple(int(x) for x in rt)
    assert np.array_equal(h.band(), ref.astype(np.float32))


def test_compact_wire_format_edges(c_oracle):
    from gretel_b200 import util
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    rng = np.random.default_rng(77)
    for trial in range(6):
        N = int(rng.integers(2, 60))
        rank, off, codes = synth.random_packed(rng, N, int(rng.integers(0, 400)), int(rng.integers(2, 12)), p_special=0.3)
        W = N + 1
        klen, codes4, n_codes = util.compact_packed(off, codes)
        ref, rt = c_oracle.ingest(rank, off, codes, N, W)
        h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
        assert h.ingest_packed_compact(rank, klen, codes4, n_codes) == tuple(int(x) for x in rt)
        assert np.array_equal(h.band(), ref.astype(np.float32))


@pytest.mark.parametrize("kernel,name,n_reads", [(5, "hiv", 20_000), (5, "metagenome", 60_000), (4, "hiv", 20_000)])
def test_bitsliced_kernels_repeatable_under_stress(c_oracle, kernel, name, n_reads):
    """The warp-specialised kernel hands bit-planes from builder warps to counting warps through full/empty
    mbarriers (compute-sanitizer's racecheck does not model those and flags every such hand-over, see
    profiles/r1_sanitizer.txt): 200 back-to-back launches must all give the oracle's matrix bit for bit."""
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    w = synth.scaled(synth.WORKLOADS[name], n_reads)
    d = synth.generate(w)
    N, W = w.n_snps, d["max_k"] - 1
    ref, rt = c_oracle.ingest(d["rank"], d["off"], d["codes"], N, W)
    ref = ref.astype(np.float32)
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
    h.set_ingest_kernel(kernel)
    reps = 200
    assert reps * float(ref.max()) < 2 ** 24               # the float32 matrix stays exact
    for it in range(reps):
        totals = h.ingest_packed(d["rank"], d["off"], d["codes"])       # counts and totals accumulate
    assert totals == tuple(reps * int(x) for x in rt)
    assert np.array_equal(h.band(), reps * ref)


@pytest.mark.parametrize("name,n_reads", [("hiv", 30_000), ("metagenome", 100_000), ("ont", 300)])
@pytest.mark.parametrize("n_chunks", [1, 3])
def test_dense_wire_format(c_oracle, name, n_reads, n_chunks):
    """hx_ingest_host_dense: uint8 rank deltas, uint8/uint16 SNP counts, 2-bit alleles + exception list; one
    synchronous call, or chunks enqueued back to back (copy of chunk i+1 overlapping chunk i)."""
    from gretel_b200 import util
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    w = synth.scaled(synth.WORKLOADS[name], n_reads)
    d = synth.generate(w)
    W = d["max_k"] - 1
    ref, rt = c_oracle.ingest(d["rank"], d["off"], d["codes"], w.n_snps, W)
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, w.n_snps, band_w=W)
    if n_chunks == 1:
        dense = util.dense_packed(d["rank"], d["off"], d["codes"])
        assert dense.klen.dtype == (np.uint16 if d["max_k"] > 255 else np.uint8)
        totals = h.ingest_packed_dense(dense)
    else:
        chunks = util.dense_chunks(d["rank"], d["off"], d["codes"], n_chunks)
        assert len(chunks) == n_chunks and sum(c.n_reads for c in chunks) == len(d["rank"])
        for c in chunks:
            assert h.ingest_packed_dense(c, wait=False) is None
        totals = h.ingest_totals()
    assert totals == tuple(int(x) for x in rt)
    assert np.array_equal(h.band(), ref.astype(np.float32))


def test_dense_wire_format_edges(c_oracle):
    """Empty input, reads of 0/1 SNPs, many N/-/_ alleles, rank gaps >= 255 (escapes), lengths that are not a
    multiple of the 16-read / 16-allele decode granules, repeated calls on one matrix."""
    from gretel_b200 import util
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    rng = np.random.default_rng(78)
    escapes = 0
    for trial in range(8):
        N = int(rng.integers(2, 60)) if trial < 5 else int(rng.integers(2000, 6000))
        n_reads = (int(rng.integers(0, 400)) if trial else 0) if trial < 5 else int(rng.integers(8, 20))
        rank, off, codes = synth.random_packed(rng, N, n_reads, int(rng.integers(2, 12)), p_special=0.3)
        W = min(N + 1, 12)
        dense = util.dense_packed(rank, off, codes)
        assert len(dense.esc_idx) == int((np.diff(rank.astype(np.int64), prepend=0) >= 255).sum())
        escapes += len(dense.esc_idx)                           # sparse reads over thousands of sites
        ref, rt = c_oracle.ingest(rank, off, codes, N, W)
        h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
        assert h.ingest_packed_dense(dense) == tuple(int(x) for x in rt)
        assert np.array_equal(h.band(), ref.astype(np.float32))
        # a second, asynchronous pass over the same reads doubles every count
        h2 = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
        for c in util.dense_chunks(rank, off, codes, 2) * 2:
            h2.ingest_packed_dense(c, wait=False)
        assert h2.ingest_totals() == tuple(2 * int(x) for x in rt)
        assert np.array_equal(h2.band(), 2 * ref.astype(np.float32))
    assert escapes > 0


def test_packed_exchange_format_single_gpu(c_oracle):
    """hx_counts_pack / hx_counts_unpack (the multi-GPU exchange with uint16 lanes): doubling the packed words
    stands in for a 2-rank sum all-reduce; an input that could overflow a lane is refused untouched."""
    import torch
    from gretel_b200.dist import _DevBuf
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    rng = np.random.default_rng(5)
    N = 300
    rank, off, codes = synth.random_packed(rng, N, 20_000, 14, p_special=0.1)
    W = int(np.diff(off).max()) - 1
    ref, rt = c_oracle.ingest(rank, off, codes, N, W)
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
    h.ingest_packed(rank, off, codes)
    ptr, n = h.counts_pack(2)
    assert n == (((N + 2) * W * 25 + 3) & ~3) + 4
    dev = torch.device("cuda", h.device)
    packed = torch.as_tensor(_DevBuf(ptr, n, "<i4"), device=dev)
    with torch.cuda.stream(torch.cuda.ExternalStream(h.stream, device=dev)):
        packed.add_(packed)
    assert h.counts_unpack()
    assert np.array_equal(h.band(), 2 * ref.astype(np.float32))
    # the same counts cannot be summed over 20000 ranks in 16-bit lanes: refused, counts left as they were
    h2 = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
    h2.ingest_packed(rank, off, codes)
    assert ref.max() > 65535 // 20000
    ptr, n = h2.counts_pack(20000)
    packed = torch.as_tensor(_DevBuf(ptr, n, "<i4"), device=dev)
    with torch.cuda.stream(torch.cuda.ExternalStream(h2.stream, device=dev)):
        packed.add_(packed)
    assert not h2.counts_unpack()
    assert np.array_equal(h2.band(), ref.astype(np.float32))


def test_load_from_packed_host_paths_agree(monkeypatch):
    """util.load_from_packed from host arrays: allele bytes as they are + encoded ranks / SNP counts (default, "slim"),
    the host-side dense encoder pipelined in C (HX_HOST_PIPELINE=dense), the packed arrays in chunks (=packed), the
    Python-driven dense chunks and one plain copy (=off) all give the same matrix."""
    from gretel_b200 import util
    d = synth.generate(synth.scaled(synth.WORKLOADS["metagenome"], 250_000))
    N, W = d["n_snps"], d["max_k"] - 1
    assert len(d["rank"]) >= util.DENSE_MIN_READS
    a = util.load_from_packed(d["rank"], d["off"], d["codes"], N, band_w=W)
    ref = a.band()
    monkeypatch.setenv("HX_HOST_PIPELINE", "dense")
    b = util.load_from_packed(d["rank"], d["off"], d["codes"], N, band_w=W)
    monkeypatch.setenv("HX_HOST_PIPELINE", "off")
    c = util.load_from_packed(d["rank"], d["off"], d["codes"], N, band_w=W)
    assert b.launch_count() > c.launch_count() and a.launch_count() > c.launch_count()    # the decode kernels ran
    monkeypatch.setenv("HX_HOST_PIPELINE", "packed")
    f = util.load_from_packed(d["rank"], d["off"], d["codes"], N, band_w=W)
    monkeypatch.delenv("HX_HOST_PIPELINE")
    e = util.load_from_packed(d["rank"], d["off"], d["codes"], N, band_w=W, wire="dense")
    for x in (b, c, e, f):
        assert (a.n_slices, a.n_crumbs, a.L) == (x.n_slices, x.n_crumbs, x.L)
        assert np.array_equal(ref, x.band())
    # unsorted input cannot be dense-encoded: the library notices and ships it as it is
    perm = np.random.default_rng(0).permutation(len(d["rank"]))
    k = np.diff(d["off"])
    off2 = np.concatenate([[0], np.cumsum(k[perm])]).astype(np.int64)
    idx = np.repeat(d["off"][:-1][perm], k[perm]) + (np.arange(int(off2[-1])) - np.repeat(off2[:-1], k[perm]))
    for mode in ("dense", "slim"):
        monkeypatch.setenv("HX_HOST_PIPELINE", mode)
        u = util.load_from_packed(d["rank"][perm], off2, d["codes"][idx], N, band_w=W)
        assert (a.n_slices, a.n_crumbs, a.L) == (u.n_slices, u.n_crumbs, u.L)
        assert np.array_equal(ref, u.band())
    # sorted at first, shuffled from the middle on: the chunks that were sorted go out encoded, the rest as it is
    half = len(perm) // 2
    perm2 = np.concatenate([np.arange(half), half + np.random.default_rng(1).permutation(len(perm) - half)])
    off3 = np.concatenate([[0], np.cumsum(k[perm2])]).astype(np.int64)
    idx3 = np.repeat(d["off"][:-1][perm2], k[perm2]) + (np.arange(int(off3[-1])) - np.repeat(off3[:-1], k[perm2]))
    monkeypatch.delenv("HX_HOST_PIPELINE")
    v = util.load_from_packed(d["rank"][perm2], off3, d["codes"][idx3], N, band_w=W)
    assert (a.n_slices, a.n_crumbs, a.L) == (v.n_slices, v.n_crumbs, v.L)
    assert np.array_equal(ref, v.band())


def test_dense_wire_format_rejects_bad_input():
    from gretel_b200 import util, _lib
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    with pytest.raises(ValueError):
        util.dense_packed(np.array([3, 1], np.int32), np.array([0, 2, 4], np.int64), np.zeros(4, np.uint8))
    # an exception index past the allele stream, and a 2-bit field of 3 at an exception (code 7)
    rank, off = np.array([0, 1], np.int32), np.array([0, 3, 6], np.int64)
    dense = util.dense_packed(rank, off, np.array([0, 1, 5, 2, 3, 0], np.uint8))
    for bad in ("pos", "field", "n_codes", "klen", "escape"):
        d2 = util.DensePacked(dense.rank_delta.copy(), dense.esc_idx, dense.esc_delta, dense.klen.copy(),
                              dense.codes2.copy(), dense.exc_pos.copy(), dense.n_reads, dense.n_codes)
        if bad == "pos":
            d2.exc_pos[0] = 600
        elif bad == "field":
            d2.codes2[0] |= 3 << 4                               # allele 2 (the exception) -> field 3 -> code 7
        elif bad == "n_codes":
            d2.n_codes = 5                                       # SNP counts add up to 6
            d2.exc_pos = d2.exc_pos[:0]
        elif bad == "klen":
            d2.klen[1] = 200                                     # offsets would run past the alleles shipped
        else:                                                    # a negative escape delta: ranks would decrease
            d2.rank_delta[1] = 255
            d2.esc_idx, d2.esc_delta = np.array([1], np.int64), np.array([-7], np.int32)
        h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, 6, band_w=4)
        if bad == "escape":
            h.ingest_packed_dense(d2)                            # clamped to "same rank": still a valid input
            continue
        with pytest.raises(_lib.HanselxError):
            h.ingest_packed_dense(d2)
        h.close()
    # the library is still usable afterwards (no sticky CUDA error)
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, 6, band_w=4)
    assert h.ingest_packed_dense(dense)[0] == 2


@pytest.mark.parametrize("shape", [(600, 40, 300_000, 150), (600, 100, 120_000, 200), (3000, 300, 400_000, 150)],
                         ids=["7k-reads-per-rank", "wide-reads-deep", "1k-reads-per-rank"])
@pytest.mark.parametrize("kernel", [0, 4, 5, 3, 6, 7])
def test_deep_coverage_runs_bit_exact(c_oracle, shape, kernel):
    """Runs of thousands of reads per rank (several 1024-read batches per run, several CTAs per
    run) - the regime of the full-size configs - at a size the C oracle checks in a second."""
    G, N, R, L = shape
    w = synth.Workload("deep", 1, G, N, R, L, 0.01, 0.004)
    d = synth.generate(w, seed=5)
    W = d["max_k"] - 1
    ref, rt = c_oracle.ingest(d["rank"], d["off"], d["codes"], N, W)
    band, totals = _gpu_band(d["rank"], d["off"], d["codes"], N, W, kernel)
    assert totals == tuple(int(x) for x in rt)
    assert np.array_equal(band, ref.astype(np.float32))


def test_adversarial_shapes(c_oracle):
    """All reads on one rank; only k=2 reads; a read spanning the whole region; isolated ranks far apart."""
    rng = np.random.default_rng(11)
    cases = []
    k = rng.integers(2, 21, size=5000)                              # one rank, 5000 reads
    off = np.concatenate([[0], np.cumsum(k)]).astype(np.int64)
    cases.append((np.full(5000, 3, np.int32), off, rng.integers(0, 6, size=off[-1]).astype(np.uint8), 40, 19))
    off = (np.arange(3001) * 2).astype(np.int64)                    # only pairs, every rank
    cases.append((np.sort(rng.integers(0, 99, size=3000)).astype(np.int32), off,
                  rng.integers(0, 4, size=6000).astype(np.uint8), 100, 1))
    off = np.array([0, 50, 52, 54], np.int64)                       # one read covers everything
    cases.append((np.array([0, 0, 48], np.int32), off, rng.integers(0, 4, size=54).astype(np.uint8), 50, 49))
    ranks = np.sort(np.repeat(np.array([0, 500, 1000, 1990], np.int32), 700))   # isolated ranks (ring jumps)
    k = rng.integers(2, 11, size=len(ranks))
    off = np.concatenate([[0], np.cumsum(k)]).astype(np.int64)
    cases.append((ranks, off, rng.integers(0, 5, size=off[-1]).astype(np.uint8), 2000, 9))
    for rank, off, codes, N, W in cases:
        ref, rt = c_oracle.ingest(rank, off, codes, N, W)
        for kernel in (0, 1, 3, 4, 5, 6, 7):
            band, totals = _gpu_band(rank, off, codes, N, W, kernel)
            assert totals == tuple(int(x) for x in rt), (N, kernel)
            assert np.array_equal(band, ref.astype(np.float32)), (N, kernel)


_CONFIG3 = {}


def _config3(c_oracle):
    """BASELINE.json configs[2] at full size, generated and ingested by the C oracle once per session."""
    if not _CONFIG3:
        d = synth.generate(synth.WORKLOADS["metagenome"])
        N, W = d["n_snps"], d["max_k"] - 1
        ref, rt = c_oracle.ingest(d["rank"], d["off"], d["codes"], N, W)
        _CONFIG3.update(d=d, N=N, W=W, ref=ref, rt=rt)
    return _CONFIG3


def test_full_size_config3_bit_exact(c_oracle):
    """BASELINE.json configs[2] at FULL size (10M x 150 bp reads, 10k SNPs, 1.1 G observations): the GPU band
    equals the C oracle's bit for bit, and so do the totals (n_slices, n_crumbs, covered SNPs, sentinels)."""
    c3 = _config3(c_oracle)
    d, N, W, ref, rt = c3["d"], c3["N"], c3["W"], c3["ref"], c3["rt"]
    for kernel in (0, 2):                  # auto = the tensor-core kernel at this depth; 2 = bit-sliced POPC kernel
        band, totals = _gpu_band(d["rank"], d["off"], d["codes"], N, W, kernel)
        assert totals == tuple(int(x) for x in rt)
        assert totals[1] > 1_000_000_000
        assert np.array_equal(band, ref.astype(np.float32))
    # ... and the first haplotypes recovered from that full-size matrix (10k sites, L = 15)
    from gretel_b200 import util
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
    h.load_band(band)
    util.set_totals(h, totals[0], totals[1], totals[2])
    assert h.L == 15
    orig = h.copy()
    cur = ref.astype(np.float32)
    cur0 = cur.copy()
    for _ in range(3):
        pc, res = c_oracle.generate_path(cur, cur0, N, W, h.L)
        got = h.generate_path_codes(orig)
        assert pc is not None and got[0] is not None
        assert np.array_equal(got[0], pc)
        for a, b in zip(got[1:], res):
            assert a == pytest.approx(b, rel=1e-6)
        ratio = max(res[2], 0.01)
        assert h.reweight_path_codes(got[0], ratio) == pytest.approx(c_oracle.reweight_path(cur, N, W, pc, ratio), rel=1e-9)
    assert np.array_equal(h.band(), cur)


@pytest.mark.parametrize("L", range(1, 9))
def test_config5_recovery_sweep_parity(c_oracle, L):
    """BASELINE.json configs[4]: the 10k-SNP matrix of configs[2] at full size, up to 50 ranked haplotypes at
    lookback L (the sweep bench.py times), through the resident driver loop hx_recover (gretel/cmd.py:148-161):
    every haplotype identical to the C oracle's, every scalar within 1e-6 relative, the reweighted matrix equal."""
    from gretel_b200 import util
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    c3 = _config3(c_oracle)
    N, W, ref, rt = c3["N"], c3["W"], c3["ref"], c3["rt"]
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
    h.load_band(ref.astype(np.float32))
    util.set_totals(h, int(rt[0]), int(rt[1]), int(rt[2]))
    orig = h.copy()
    h.L = L
    paths, stats = h.recover_codes(orig, 50, 0.01)
    cur = ref.astype(np.float32)
    cur0 = cur.copy()
    n = 0
    for it in range(50):
        pc, res = c_oracle.generate_path(cur, cur0, N, W, L)
        if pc is None:
            break
        ratio = max(res[2], 0.01)
        removed = c_oracle.reweight_path(cur, N, W, pc, ratio)
        assert it < len(paths), "the GPU loop stopped early at haplotype %d" % it
        assert np.array_equal(paths[it], pc), "L=%d haplotype %d differs from the oracle" % (L, it)
        for got, exp in zip(stats[it], (res[0], res[1], res[2], ratio, removed)):
            assert got == pytest.approx(exp, rel=1e-6)
        n += 1
    assert len(paths) == n and n >= 10
    assert np.array_equal(h.band(), cur)
    h.close()
    orig.close()


_CONFIG4 = {}


@pytest.mark.parametrize("kernel", [0, 3], ids=["auto-tensor-core", "bit-plane-tiles"])
def test_config4_full_size_bit_exact(c_oracle, kernel):
    """BASELINE.json configs[3] at FULL size (100k ONT-like reads, ~300 SNPs/read, 4.7 G observations), bit for bit
    against the C oracle: through the long-read tensor-core kernel (what auto picks) - into a cleared matrix (plain
    stores) and again on top of it (read-modify-write) - and through the cp.async-staged bit-plane tile kernel."""
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    if not _CONFIG4:
        d = synth.generate(synth.WORKLOADS["ont"])
        N, W = d["n_snps"], d["max_k"] - 1
        ref, rt = c_oracle.ingest(d["rank"], d["off"], d["codes"], N, W)
        _CONFIG4.update(d=d, N=N, W=W, ref=ref.astype(np.float32), rt=tuple(int(x) for x in rt))
    d, N, W, ref, rt = (_CONFIG4[k] for k in ("d", "N", "W", "ref", "rt"))
    assert len(d["rank"]) >= 6 * N                      # dense enough for the staged tiles
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
    h.set_ingest_kernel(kernel)
    totals = h.ingest_packed(d["rank"], d["off"], d["codes"])
    assert totals == rt
    assert totals[1] > 4_000_000_000
    if kernel == 0:
        totals2 = h.ingest_packed(d["rank"], d["off"], d["codes"])
        assert totals2 == tuple(2 * x for x in rt)
        ref = ref * 2
    band = h.band()
    h.close()
    assert np.array_equal(band, ref)


@pytest.mark.parametrize("seed", range(4))
def test_parity_probe_counts_what_the_oracle_counts(c_oracle, seed):
    """hx_probe_expected_rows (the independent per-row recount bench.py's parity_probe relies on) equals the row
    sums of the C oracle's band, and hx_counts_row_sums equals the row sums of what the GPU ingested."""
    import torch
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    rng = np.random.default_rng(800 + seed)
    N = int(rng.integers(3, 300))
    rank, off, codes = synth.random_packed(rng, N, int(rng.integers(10, 4000)), int(rng.integers(2, 30)),
                                           p_special=0.2)
    W = max(1, int(np.diff(off).max()) - 1)
    ref, rt = c_oracle.ingest(rank, off, codes, N, W)
    dev = torch.device("cuda", 0)
    t_rank, t_off, t_codes = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (rank, off, codes))
    if t_codes.numel() == 0:
        t_codes = torch.zeros(16, dtype=torch.uint8, device=dev)
    rows_exp = torch.zeros(N + 2, dtype=torch.int64, device=dev)
    rows_got = torch.zeros(N + 2, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
    h.ingest_device(t_rank.data_ptr(), t_off.data_ptr(), t_codes.data_ptr(), len(rank))
    h.probe_expected_rows(t_rank.data_ptr(), t_off.data_ptr(), t_codes.data_ptr(), len(rank), rows_exp.data_ptr())
    h.counts_row_sums(rows_got.data_ptr())
    h.sync()
    want = ref.reshape(N + 2, -1).sum(axis=1).astype(np.int64)
    assert np.array_equal(rows_exp.cpu().numpy(), want)
    assert np.array_equal(rows_got.cpu().numpy(), want)
    assert int(want.sum()) == int(rt[1]) + int(rt[3])
    h.close()


def test_tensor_core_kernel_deep_single_rank(c_oracle):
    """300k reads on ONE rank (every count far above what an 8-bit or 16-bit lane could hold, and thousands of
    MMAs accumulating into one TMEM stage from all expander warps), plus a neighbouring rank: the int32
    accumulation of the tensor-core kernel is exact."""
    rng = np.random.default_rng(21)
    k = rng.integers(20, 31, size=300_000)
    off = np.concatenate([[0], np.cumsum(k)]).astype(np.int64)
    rank = np.where(np.arange(len(k)) < 290_000, 5, 6).astype(np.int32)
    codes = rng.choice(np.array([0, 1, 2, 3, 4, 5], np.uint8), size=int(off[-1]), p=[0.5, 0.3, 0.1, 0.08, 0.01, 0.01])
    N, W = 40, 29
    ref, rt = c_oracle.ingest(rank, off, codes, N, W)
    band, totals = _gpu_band(rank, off, codes, N, W, 6)
    assert totals == tuple(int(x) for x in rt)
    assert ref.max() > 60_000
    assert np.array_equal(band, ref.astype(np.float32))


def test_unsorted_short_reads_take_the_counting_sort_kernel(c_oracle, monkeypatch):
    """Short reads that are not sorted by rank: the first such ingestion of a process falls back to one RED per pair;
    the library remembers, and from then on queues the counting-sort tensor-core kernel as the fallback.  Both give the
    oracle's counts; sorted input afterwards still goes through the sorted-run kernels (with the fallback idle)."""
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    monkeypatch.setenv("HX_HOST_PIPELINE", "off")          # one launch per ingest_packed: kernel_ms covers all of it
    rng = np.random.default_rng(99)
    N, R, mk = 3000, 400_000, 24
    rank, off, codes = synth.random_packed(rng, N, R, mk, p_special=0.05, sort=False)
    W = mk - 1
    ref, rt = c_oracle.ingest(rank, off, codes, N, W)
    ms = []
    for _ in range(3):
        h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W)
        totals = h.ingest_packed(rank, off, codes)
        ms.append(h.kernel_ms("ingest"))
        assert totals == tuple(int(x) for x in rt)
        assert np.array_equal(h.band(), ref.astype(np.float32))
        h.ingest_packed(rank, off, codes)                  # on top of existing counts
        assert np.array_equal(h.band(), 2 * ref.astype(np.float32))
        h.close()
    print("unsorted short reads: %.3f ms (first), %.3f ms, %.3f ms" % tuple(ms))
    order = np.argsort(rank, kind="stable")
    k = np.diff(off)
    off_s = np.concatenate([[0], np.cumsum(k[order])]).astype(np.int64)
    idx = np.repeat(off[:-1][order], k[order]) + (np.arange(int(off_s[-1])) - np.repeat(off_s[:-1], k[order]))
    band, totals = _gpu_band(rank[order], off_s, codes[idx], N, W)
    assert totals == tuple(int(x) for x in rt)
    assert np.array_equal(band, ref.astype(np.float32))


def _packed_from(rank, k, rng, p_special=0.02):
    order = np.argsort(rank, kind="stable")
    rank, k = rank[order].astype(np.int32), k[order]
    off = np.concatenate([[0], np.cumsum(k)]).astype(np.int64)
    codes = rng.integers(0, 4, size=int(off[-1])).astype(np.uint8)
    sp = rng.random(len(codes)) < p_special
    codes[sp] = rng.integers(4, 7, size=int(sp.sum())).astype(np.uint8)
    return rank, off, codes


@pytest.mark.parametrize("shape", ["one_run", "many_runs_per_cta", "mostly_narrow_reads", "rank_jumps", "dense_then_sparse"])
@pytest.mark.parametrize("kernel", [6, 2, 5])
def test_slices_and_job_tables_on_skewed_shapes(c_oracle, shape, kernel):
    """The sorted-run kernels cut the reads into CTA slices of equal WEIGHT (alleles, reads, ranks) with a search over
    off[] / rank[], and the tensor-core kernel builds per-CTA job tables from the run list in batches of 256 runs:
    shapes that put the cuts inside one giant run, that give a CTA more than 256 runs, whose weight is almost all
    per-read or per-rank, and whose density changes abruptly."""
    rng = np.random.default_rng({"one_run": 1, "many_runs_per_cta": 2, "mostly_narrow_reads": 3, "rank_jumps": 4, "dense_then_sparse": 5}[shape])
    if shape == "one_run":
        N, R = 40, 300_000
        rank = np.full(R, 5)
        k = rng.integers(2, 21, size=R)
    elif shape == "many_runs_per_cta":
        N, R = 300_000, 150_000                  # ~118k populated ranks over 148 CTAs: ~800 runs each, 4 batches of 256
        rank = rng.integers(0, N - 8, size=R)
        k = rng.integers(2, 9, size=R)
    elif shape == "mostly_narrow_reads":
        N, R = 2000, 250_000
        rank = rng.integers(0, N - 12, size=R)
        k = np.where(rng.random(R) < 0.9, rng.integers(0, 2, size=R), rng.integers(2, 13, size=R))
    elif shape == "rank_jumps":
        N, R = 50_000, 120_000
        rank = rng.choice(np.arange(0, N - 32, 977), size=R)      # 52 populated ranks, 977 apart
        k = rng.integers(2, 31, size=R)
    else:
        N, R = 6000, 200_000
        dense = rng.random(R) < 0.8
        rank = np.where(dense, rng.integers(0, 300, size=R), rng.integers(300, N - 30, size=R))
        k = np.where(dense, rng.integers(20, 31, size=R), rng.integers(2, 6, size=R))
    k = np.minimum(k, N - rank)
    rank, off, codes = _packed_from(rank, k, rng)
    W = int(max(1, k.max() - 1))
    ref, rt = c_oracle.ingest(rank, off, codes, N, W)
    band, totals = _gpu_band(rank, off, codes, N, W, kernel)
    assert totals == tuple(int(x) for x in rt)
    assert np.array_equal(band, ref.astype(np.float32))
