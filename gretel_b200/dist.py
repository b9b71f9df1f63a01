"""Multi-GPU ingestion: reads shard naturally (the matrix is a sum over reads), so each
GPU builds a partial integer band from its slice of the rank-sorted reads and the
partials are summed with one integer all-reduce (NCCL over NVLink; gloo in CPU tests).
This replaces the reference's fork-shared matrix written by ``n_threads`` window workers
(gretel/util.py:294-326).  Recovery is sequential across haplotypes and runs on one GPU
(every rank holds the full matrix after the all-reduce; rank 0 recovers).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(off, world_size):
    """Split reads into ``world_size`` contiguous chunks balanced by pair count
    sum k(k-1)/2 (not by read count).  Returns int64[world_size+1] read indices."""
    off = np.asarray(off, dtype=np.int64)
    k = np.diff(off)
    work = np.concatenate([[0], np.cumsum(k * (k - 1) // 2)])
    total = work[-1]
    targets = (total * np.arange(1, world_size)) // world_size
    cuts = np.searchsorted(work, targets, side="left")
    return np.concatenate([[0], cuts, [len(k)]]).astype(np.int64)


def take_shard(rank, off, codes, lo, hi):
    """Slice reads [lo,hi) without copying codes (offsets stay absolute)."""
    return rank[lo:hi], off[lo:hi + 1], codes


class _DevBuf:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def reduce_counts(hansel, dst=0, group=None):
    """Sum the pending integer counts and totals onto rank ``dst`` only (recovery runs on one GPU, so the
    other ranks do not need the full matrix): about half the traffic of the all-reduce."""
    import torch
    import torch.distributed as dist
    cptr, cn, tptr, tn = hansel.counts_buffer()
    dev = torch.device("cuda", hansel.device)
    with torch.cuda.device(dev):
        s = torch.cuda.ExternalStream(hansel.stream, device=dev)
        with torch.cuda.stream(s):
            counts = torch.as_tensor(_DevBuf(cptr, cn, "<i4"), device=dev)
            totals = torch.as_tensor(_DevBuf(tptr, tn, "<i8"), device=dev)
            dist.reduce(counts, dst=dst, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(totals, op=dist.ReduceOp.SUM, group=group)


def allreduce_counts_packed(hansel, group=None):
    """allreduce_counts with half the bytes: counts travel as uint16 lanes packed into uint32 words
    (hx_counts_pack); falls back to the plain all-reduce when a lane could overflow on some rank.
    Returns True if the packed exchange was used."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    _c, _n, tptr, tn = hansel.counts_buffer()
    pptr, pn = hansel.counts_pack(world)
    dev = torch.device("cuda", hansel.device)
    with torch.cuda.device(dev):
        s = torch.cuda.ExternalStream(hansel.stream, device=dev)
        with torch.cuda.stream(s):
            packed = torch.as_tensor(_DevBuf(pptr, pn, "<i4"), device=dev)
            totals = torch.as_tensor(_DevBuf(tptr, tn, "<i8"), device=dev)
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(totals, op=dist.ReduceOp.SUM, group=group)
    if hansel.counts_unpack():
        return True
    cptr, cn, _t, _tn = hansel.counts_buffer()
    with torch.cuda.device(dev):
        s = torch.cuda.ExternalStream(hansel.stream, device=dev)
        with torch.cuda.stream(s):
            counts = torch.as_tensor(_DevBuf(cptr, cn, "<i4"), device=dev)
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return False


def allreduce_counts(hansel, group=None):
    """Sum the pending integer counts and totals of ``hansel`` across ranks, in place."""
    import torch
    import torch.distributed as dist
    cptr, cn, tptr, tn = hansel.counts_buffer()
    dev = torch.device("cuda", hansel.device)
    with torch.cuda.device(dev):
        s = torch.cuda.ExternalStream(hansel.stream, device=dev)
        with torch.cuda.stream(s):
            counts = torch.as_tensor(_DevBuf(cptr, cn, "<i4"), device=dev)      # uint32 sums are exact mod 2^32
            totals = torch.as_tensor(_DevBuf(tptr, tn, "<i8"), device=dev)
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(totals, op=dist.ReduceOp.SUM, group=group)


class OverlappedAllreduce:
    """Back-to-back ingestion jobs (one matrix per region / sample): the all-reduce of job i runs on a
    communication stream while the pair expansion of job i+1 runs on the compute stream, with two matrices taking
    turns.  Every job still does all of its work (zero, expand, sum across GPUs); only the exposure of the
    collective changes.  The ingestion kernels are persistent (one CTA per SM), so a few SMs are left to NCCL
    (``free_sms``)."""

    def __init__(self, hansels, group=None, free_sms=8, packed=False):
        """``packed``: the counts travel as uint16 lanes packed into uint32 words (hx_counts_pack: 25 words instead of
        49 per cell); pack and write-back are kernels on the compute / communication stream, nothing waits for the
        host.  Safe while every count * world <= 65535 (``packing_is_safe``); the pack kernel flags anything larger,
        the flag is summed with the data, the write-back then leaves the partials alone and ``drain`` raises."""
        import torch
        import torch.distributed as dist
        self.hs, self.group, self.dist, self.torch = list(hansels), group, dist, torch
        self.packed = bool(packed)
        self.world = dist.get_world_size(group)
        self.pbufs = [None, None]
        assert len(self.hs) == 2
        dev = torch.device("cuda", self.hs[0].device)
        self.dev = dev
        self.main = torch.cuda.Stream(device=dev)
        self.comm = torch.cuda.Stream(device=dev)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        self.bufs = []
        for h in self.hs:
            h.sync()
            h.set_stream(self.main.cuda_stream)
            h.set_ingest_sms(max(1, sms - free_sms))
            cptr, cn, tptr, tn = h.counts_buffer()
            self.bufs.append((torch.as_tensor(_DevBuf(cptr, cn, "<i4"), device=dev),
                              torch.as_tensor(_DevBuf(tptr, tn, "<i8"), device=dev)))
        self.ingested = [torch.cuda.Event() for _ in range(2)]
        self.reduced = [torch.cuda.Event() for _ in range(2)]
        self.used = [False, False]
        self.i = 0

    def step(self, ingest):
        """``ingest(h)`` enqueues one job's reset + pair expansion on matrix ``h`` (its stream is ``self.main``)."""
        b = self.i & 1
        h = self.hs[b]
        if self.used[b]:
            self.main.wait_event(self.reduced[b])            # the matrix's previous job has been summed and consumed
        ingest(h)
        counts, totals = self.bufs[b]
        if self.packed:
            pptr, pn = h.counts_pack(self.world)               # on the compute stream, behind the expansion
            if self.pbufs[b] is None or self.pbufs[b][0] != (pptr, pn):
                self.pbufs[b] = ((pptr, pn), self.torch.as_tensor(_DevBuf(pptr, pn, "<i4"), device=self.dev))
            counts = self.pbufs[b][1]
        self.ingested[b].record(self.main)
        self.comm.wait_event(self.ingested[b])
        with self.torch.cuda.stream(self.comm):
            self.dist.all_reduce(counts, op=self.dist.ReduceOp.SUM, group=self.group)
            self.dist.all_reduce(totals, op=self.dist.ReduceOp.SUM, group=self.group)
            if self.packed:
                h.counts_unpack_async(self.comm.cuda_stream)   # sums back into the counts, still off the compute stream
            self.reduced[b].record(self.comm)
        self.used[b] = True
        self.i += 1
        return h

    def drain(self):
        """Order the compute stream behind every outstanding all-reduce."""
        for b in range(2):
            if self.used[b]:
                self.main.wait_event(self.reduced[b])
        if self.packed:
            for b in range(2):
                if self.used[b] and self.hs[b].counts_pack_overflowed():
                    raise RuntimeError("packed exchange: a count exceeded 65535 / world on some rank; the sums of the "
                                       "jobs since the last drain are partial - use packed=False")

    @staticmethod
    def packing_is_safe(hansel, group=None, headroom=2):
        """True on every rank iff the largest pending count on any rank, times world (times ``headroom`` for jobs that
        are not identical to this one), fits a uint16 lane.  Synchronises; call it once, on a representative job."""
        import torch
        import torch.distributed as dist
        world = dist.get_world_size(group)
        m = torch.tensor([hansel.counts_max()], dtype=torch.int64, device=torch.device("cuda", hansel.device))
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
        return int(m.item()) * world * headroom <= 65535


class SeamExchange:
    """Strong scaling: the rank-sorted reads are cut into ``world`` contiguous chunks (shard_bounds), so the partial
    matrices of neighbouring GPUs overlap only in the band rows their boundary reads share (a read of rank r
    touches rows r+2 .. r+k): the "seam" of at most W+1 rows behind the last rank of a chunk.  Instead of
    all-reducing the whole band, GPU g sends its seam rows to GPU g+1, which adds them; then every GPU holds
    the final values of the rows it owns ([first rank of its chunk + 2, last rank + 2), rank 0 from row 0, the last
    GPU to row N+1) and sends them to ``dst`` (recovery runs on one GPU).  Bytes on the wire: (world-1) seams of
    (W+1)*W*196 B plus each row once, instead of 2x the whole band per GPU.

    ``last_rank``: the rank value of this GPU's last read (-1 if it has no reads).  Falls back to the plain
    all-reduce (``fallback`` is True) when some chunk spans fewer ranks than a read, i.e. when rows would be shared
    by more than two neighbours."""

    def __init__(self, hansel, last_rank, group=None, dst=0, tensor_ops=None):
        import torch
        import torch.distributed as dist
        self.h, self.group, self.dist, self.torch, self.dst = hansel, group, dist, torch, dst
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        N, W = hansel.n_snps, hansel.band_w
        dev = tensor_ops["device"] if tensor_ops else torch.device("cuda", hansel.device)
        t = torch.tensor([int(last_rank)], device=dev, dtype=torch.int64)
        allr = [torch.zeros(1, device=dev, dtype=torch.int64) for _ in range(self.world)]
        dist.all_gather(allr, t, group=group)
        last = [int(x.item()) for x in allr]
        for g in range(1, self.world):                       # a GPU without reads ends where its predecessor ended
            if last[g] < 0:
                last[g] = last[g - 1]
        if last[0] < 0:
            last[0] = 0
        self.last = last
        # rows owned by GPU g: [own_lo[g], own_hi[g]); seam sent by g to g+1: [own_hi[g], seam_hi[g])
        self.own_lo = [0] + [min(last[g - 1] + 2, N + 2) for g in range(1, self.world)]
        self.own_hi = [min(last[g] + 2, N + 2) for g in range(self.world - 1)] + [N + 2]
        self.seam_hi = [min(self.own_hi[g] + W + 1, N + 2) for g in range(self.world)]
        # exact only while a seam ends inside the next GPU's own rows (chunks at least a read wide)
        # (also: rank-0 reads in two chunks share the start sentinel's row 1; a chunk without reads would have to
        # forward its predecessor's seam)
        self.fallback = any(self.seam_hi[g] > self.own_hi[g + 1] and g + 1 < self.world - 1 for g in range(self.world - 1)) \
            or any(self.own_hi[g] <= self.own_lo[g] for g in range(self.world)) \
            or any(last[g] == 0 for g in range(self.world - 1))
        self.row = W * 49

    def run(self, counts, totals):
        """``counts``: this GPU's partial band as an int32 tensor [(N+2)*W*49]; ``totals``: int64[4].  Enqueues the
        exchange on the current stream; afterwards ``dst`` holds the complete band, every rank the global totals."""
        dist, g, w, row = self.dist, self.rank, self.world, self.row
        dist.all_reduce(totals, op=dist.ReduceOp.SUM, group=self.group)
        if self.fallback:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=self.group)
            return
        ops = []
        recv_buf = None
        if g + 1 < w and self.seam_hi[g] > self.own_hi[g]:
            ops.append(dist.P2POp(dist.isend, counts[self.own_hi[g] * row:self.seam_hi[g] * row], g + 1, group=self.group))
        if g > 0 and self.seam_hi[g - 1] > self.own_hi[g - 1]:
            recv_buf = self.torch.empty((self.seam_hi[g - 1] - self.own_hi[g - 1]) * row, dtype=counts.dtype, device=counts.device)
            ops.append(dist.P2POp(dist.irecv, recv_buf, g - 1, group=self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        if recv_buf is not None:
            counts[self.own_hi[g - 1] * row:self.seam_hi[g - 1] * row] += recv_buf
        # owned rows -> dst
        ops = []
        if g == self.dst:
            for src in range(w):
                if src != g and self.own_hi[src] > self.own_lo[src]:
                    ops.append(dist.P2POp(dist.irecv, counts[self.own_lo[src] * row:self.own_hi[src] * row], src, group=self.group))
        elif self.own_hi[g] > self.own_lo[g]:
            ops.append(dist.P2POp(dist.isend, counts[self.own_lo[g] * row:self.own_hi[g] * row], self.dst, group=self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()

    def bytes_on_wire(self):
        if self.fallback:
            return None
        seams = sum((self.seam_hi[g] - self.own_hi[g]) for g in range(self.world - 1)) * self.row * 4
        rows = sum((self.own_hi[g] - self.own_lo[g]) for g in range(self.world) if g != self.dst) * self.row * 4
        return seams + rows


def seam_exchange_counts(hansel, last_rank, group=None, dst=0, plan=None):
    """Run the seam exchange on the pending integer counts of ``hansel`` (device tensors over the C-ABI buffers)."""
    import torch
    cptr, cn, tptr, tn = hansel.counts_buffer()
    dev = torch.device("cuda", hansel.device)
    with torch.cuda.device(dev):
        s = torch.cuda.ExternalStream(hansel.stream, device=dev)
        with torch.cuda.stream(s):
            counts = torch.as_tensor(_DevBuf(cptr, cn, "<i4"), device=dev)
            totals = torch.as_tensor(_DevBuf(tptr, tn, "<i8"), device=dev)
            if plan is None:
                plan = SeamExchange(hansel, last_rank, group=group, dst=dst)
            plan.run(counts, totals)
    return plan


class PipelinedIngest:
    """Ingestion of device-resident, rank-sorted reads in ``segments`` launches with the
    all-reduce of the finished band rows overlapped with the next launch.

    After the reads [0, i) are ingested, every band row pj <= rank[i] + 1 is final on this GPU
    (a read of rank r only touches rows pj >= r + 2), so that slice of the partial matrix can
    already be summed across GPUs on a second stream while the kernel works on the next
    segment.  Only the last segment's all-reduce stays exposed."""

    def __init__(self, hansel, rank_host, n_reads, segments=4, group=None):
        import torch
        import torch.distributed as dist
        self.h, self.group, self.dist, self.torch = hansel, group, dist, torch
        S = max(1, int(segments))
        self.cuts = [(n_reads * s) // S for s in range(S + 1)]
        cell = hansel.band_w * 49
        # last finished row after each segment; must agree on every rank -> take the minimum
        # (rank-0 reads also add the start sentinel into row 1, util.py:262-266: while the cut still lies
        # inside them nothing is final yet)
        rows = [min(int(rank_host[self.cuts[s + 1]]) + 1, hansel.n_snps + 1) if int(rank_host[self.cuts[s + 1]]) > 0
                else -1 for s in range(S - 1)]
        dev = torch.device("cuda", hansel.device)
        if dist.is_initialized() and dist.get_world_size(group) > 1 and rows:
            t = torch.tensor(rows, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            rows = [int(x) for x in t.tolist()]
        rows = [max(r, -1) for r in rows] + [hansel.n_snps + 1]
        for i in range(1, len(rows)):
            rows[i] = max(rows[i], rows[i - 1])
        self.row_hi = rows
        self.cell = cell
        cptr, cn, tptr, tn = hansel.counts_buffer()
        self.counts = torch.as_tensor(_DevBuf(cptr, cn, "<i4"), device=dev)
        self.totals = torch.as_tensor(_DevBuf(tptr, tn, "<i8"), device=dev)
        self.main = torch.cuda.ExternalStream(hansel.stream, device=dev)
        self.comm = torch.cuda.Stream(device=dev)
        self.events = [torch.cuda.Event() for _ in range(S)]
        self.done = torch.cuda.Event()

    def run(self, rank_ptr, off_ptr, codes_ptr):
        """Enqueue everything; returns without synchronising (work is ordered on hansel.stream)."""
        torch, dist = self.torch, self.dist
        h = self.h
        self.comm.wait_stream(self.main)                 # counts were zeroed on the main stream
        lo_row = 0
        for s in range(len(self.cuts) - 1):
            a, b = self.cuts[s], self.cuts[s + 1]
            if b > a:
                h.ingest_device(rank_ptr + 4 * a, off_ptr + 8 * a, codes_ptr, b - a)
            self.events[s].record(self.main)
            hi_row = self.row_hi[s]
            if hi_row >= lo_row:
                self.comm.wait_event(self.events[s])
                with torch.cuda.stream(self.comm):
                    dist.all_reduce(self.counts[lo_row * self.cell:(hi_row + 1) * self.cell],
                                    op=dist.ReduceOp.SUM, group=self.group)
                lo_row = hi_row + 1
        with torch.cuda.stream(self.comm):
            dist.all_reduce(self.totals, op=dist.ReduceOp.SUM, group=self.group)
            self.done.record(self.comm)
        self.main.wait_event(self.done)


class FusedExchange:
    """Ingestion fused with the cross-GPU sum over NVLink peer memory.

    Band rows are dealt out to the ranks in contiguous blocks; every rank maps every other rank's pending
    count buffer (CUDA IPC) and the ingestion kernels add each count straight into the GPU that owns its
    row.  When all kernels are done every rank holds the final sums of its own rows: the 68 MB all-reduce
    of the partial matrices is replaced by a barrier and an all-gather of the owned rows (1/world of the
    matrix per rank)."""

    def __init__(self, hansel, group=None):
        import torch
        import torch.distributed as dist
        self.h, self.group, self.dist, self.torch = hansel, group, dist, torch
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        dev = torch.device("cuda", hansel.device)
        mine = torch.frombuffer(bytearray(hansel.counts_ipc_export(self.world)), dtype=torch.uint8).to(dev)
        allh = [torch.empty(64, dtype=torch.uint8, device=dev) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=group)
        hansel.counts_ipc_import([bytes(t.cpu().numpy().tobytes()) for t in allh], self.rank)
        cptr, cn, tptr, tn = hansel.counts_buffer()
        self.counts = torch.as_tensor(_DevBuf(cptr, cn, "<i4"), device=dev)
        self.totals = torch.as_tensor(_DevBuf(tptr, tn, "<i8"), device=dev)
        self.own = self.counts.view(self.world, -1)[self.rank]
        self.main = torch.cuda.ExternalStream(hansel.stream, device=dev)
        self.token = torch.zeros(1, dtype=torch.int32, device=dev)
        dist.barrier(group=group)

    def reset(self):
        """Zero the local buffer; nobody may add into it before every rank has done so."""
        self.h.reset_counts()
        with self.torch.cuda.stream(self.main):
            self.dist.all_reduce(self.token, group=self.group)

    def finish(self, gather=True):
        """After the ingestion launches: wait for every rank's kernels (the totals all-reduce doubles as the
        barrier), then collect the owned rows."""
        with self.torch.cuda.stream(self.main):
            self.dist.all_reduce(self.totals, op=self.dist.ReduceOp.SUM, group=self.group)
            if gather:
                self.dist.all_gather_into_tensor(self.counts, self.own, group=self.group)

    def close(self):
        """Unmap the peers' buffers; every rank must have done so before any rank frees its own."""
        self.h.sync()
        self.dist.barrier(group=self.group)
        self.h.counts_ipc_close()
        self.dist.barrier(group=self.group)


def load_from_packed_sharded(rank, off, codes, n_snps, band_w, world_size=None, my_rank=None, device=None,
                             presharded=False, group=None):
    """Every rank calls this with the same packed reads (or, with ``presharded``, its own
    slice); returns the full Hansel on every rank."""
    import torch.distributed as dist
    from . import util
    from .hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    world_size = dist.get_world_size(group) if world_size is None else world_size
    my_rank = dist.get_rank(group) if my_rank is None else my_rank
    if not presharded:
        b = shard_bounds(off, world_size)
        rank, off, codes = take_shard(rank, off, codes, int(b[my_rank]), int(b[my_rank + 1]))
    h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, n_snps, band_w=band_w, device=device)
    h.ingest_packed(rank, off, codes)
    if world_size > 1:
        allreduce_counts(h, group=group)
    slices, crumbs, covered, _ = h.ingest_totals()
    h.finalize()
    return util.set_totals(h, slices, crumbs, covered)
