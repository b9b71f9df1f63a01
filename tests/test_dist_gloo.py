"""Host-side multi-GPU logic on CPU: world_size 2, gloo.  Each rank pair-expands its shard
(with the C oracle standing in for the GPU kernel), the partial integer bands are summed with
an all-reduce and must equal the single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest

from gretel_b200 import dist as gdist
from gretel_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_reads, out):
    import torch
    import torch.distributed as dist
    from oracle import c_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = synth.scaled(synth.WORKLOADS["hiv"], n_reads)
        d = synth.generate(w)
        W = d["max_k"] - 1
        b = gdist.shard_bounds(d["off"], world)
        r, o, c = gdist.take_shard(d["rank"], d["off"], d["codes"], int(b[rank]), int(b[rank + 1]))
        # offsets stay absolute: rebase for the oracle
        band, totals = c_oracle.ingest(r, o - o[0], c[o[0]:o[-1]], w.n_snps, W)
        t_band = torch.from_numpy(band.view(np.int32).reshape(-1))
        t_tot = torch.from_numpy(totals)
        dist.all_reduce(t_band)
        dist.all_reduce(t_tot)
        if rank == 0:
            whole, wt = c_oracle.ingest(d["rank"], d["off"], d["codes"], w.n_snps, W)
            assert np.array_equal(band, whole)
            assert np.array_equal(totals, wt)
            open(out, "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_ingest_allreduce_gloo(tmp_path, c_oracle):
    import torch.multiprocessing as mp
    out = str(tmp_path / "ok")
    mp.spawn(_worker, args=(2, _free_port(), 6000, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_bounds_balance(world):
    rng = np.random.default_rng(7)
    k = rng.integers(0, 40, size=5000)
    off = np.concatenate([[0], np.cumsum(k)]).astype(np.int64)
    b = gdist.shard_bounds(off, world)
    assert b[0] == 0 and b[-1] == len(k) and (np.diff(b) >= 0).all() and len(b) == world + 1
    work = k * (k - 1) // 2
    per = [int(work[b[i]:b[i + 1]].sum()) for i in range(world)]
    assert sum(per) == int(work.sum())
    assert max(per) - min(per) <= 2 * int(work.max()) + 1          # balanced by pairs, not by read count


def test_shard_bounds_degenerate():
    assert list(gdist.shard_bounds(np.array([0], np.int64), 4)) == [0, 0, 0, 0, 0]
    b = gdist.shard_bounds(np.array([0, 5], np.int64), 4)
    assert b[0] == 0 and b[-1] == 1


def _seam_worker(rank, world, port, case, out):
    import torch
    import torch.distributed as dist
    from oracle import c_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if case == "deep":
            w = synth.scaled(synth.WORKLOADS["hiv"], 8000)
            d = synth.generate(w)
            N = w.n_snps
        else:                                   # chunks narrower than a read / rank-0 reads in two chunks: fallback
            rng = np.random.default_rng(5)
            N = 40
            k = rng.integers(2, 21, size=3000)
            rk = np.sort(np.where(rng.random(len(k)) < 0.8, 0, rng.integers(0, N - 20, size=len(k)))).astype(np.int32)
            off = np.concatenate([[0], np.cumsum(k)]).astype(np.int64)
            d = {"rank": rk, "off": off, "codes": rng.integers(0, 6, size=int(off[-1])).astype(np.uint8), "max_k": 20}
        W = d["max_k"] - 1
        b = gdist.shard_bounds(d["off"], world)
        lo, hi = int(b[rank]), int(b[rank + 1])
        r, o, c = gdist.take_shard(d["rank"], d["off"], d["codes"], lo, hi)
        band, totals = c_oracle.ingest(r, o - o[0], c[o[0]:o[-1]], N, W)        # the oracle stands in for the GPU kernel

        class H:                                 # what SeamExchange needs of a Hansel
            n_snps, band_w, device = N, W, 0
        plan = gdist.SeamExchange(H, int(r[-1]) if len(r) else -1, tensor_ops={"device": torch.device("cpu")})
        assert plan.fallback == (case != "deep")
        t_band = torch.from_numpy(band.view(np.int32).reshape(-1))
        t_tot = torch.from_numpy(totals)
        plan.run(t_band, t_tot)
        if rank == 0:
            whole, wt = c_oracle.ingest(d["rank"], d["off"], d["codes"], N, W)
            assert np.array_equal(band, whole)
            assert np.array_equal(totals, wt)
            if case == "deep":
                assert plan.bytes_on_wire() < 0.6 * 4 * whole.size * 2          # far below an all-reduce's traffic
            open(out, "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["deep", "degenerate"])
def test_seam_exchange_gloo(tmp_path, c_oracle, world, case):
    """Strong scaling: contiguous chunks, seam rows to the right neighbour, owned rows to rank 0 == the whole matrix."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "ok")
    mp.spawn(_seam_worker, args=(world, _free_port(), case, out), nprocs=world, join=True)
    assert open(out).read() == "ok"
