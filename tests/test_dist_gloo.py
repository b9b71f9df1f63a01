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
