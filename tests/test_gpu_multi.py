"""Two-GPU parity (NCCL): sharded ingestion + (pipelined) all-reduce == single-process oracle.
Skipped on a 1-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, segments, out):
    import torch
    import torch.distributed as dist
    from gretel_b200 import dist as gdist, synth, util
    from gretel_b200.hansel import Hansel, REF_SYMBOLS, REF_UNSYMBOLS
    from oracle import c_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        if segments == 40:
            # a short gene window: most reads start before start_pos and are clamped to rank 0, so the segment
            # cuts of the pipelined all-reduce fall inside the rank-0 reads (row 1 holds their start sentinels)
            rng = np.random.default_rng(3)
            N = 60
            k = rng.integers(2, 21, size=90_000)
            rk = np.sort(np.where(rng.random(len(k)) < 0.7, 0, rng.integers(0, N - 20, size=len(k)))).astype(np.int32)
            off = np.concatenate([[0], np.cumsum(k)]).astype(np.int64)
            d = {"rank": rk, "off": off, "codes": rng.integers(0, 6, size=int(off[-1])).astype(np.uint8), "max_k": 20}
            segments = 4
        else:
            w = synth.scaled(synth.WORKLOADS["metagenome"], 150_000)
            d = synth.generate(w)
            N = w.n_snps
        W = d["max_k"] - 1
        b = gdist.shard_bounds(d["off"], world)
        lo, hi = int(b[rank]), int(b[rank + 1])
        h = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W, device=rank)
        h.counts_buffer()
        dev = torch.device("cuda", rank)
        t_rank = torch.from_numpy(d["rank"][lo:hi].copy()).to(dev)
        t_off = torch.from_numpy(d["off"][lo:hi + 1].copy()).to(dev)
        t_codes = torch.from_numpy(d["codes"]).to(dev)
        if segments == -1:                                  # fused exchange over NVLink peer memory
            fx = gdist.FusedExchange(h)
            for _ in range(2):                               # twice: the reset/barrier protocol must hold
                fx.reset()
                h.ingest_device(t_rank.data_ptr(), t_off.data_ptr(), t_codes.data_ptr(), hi - lo)
                fx.finish()
            fx.close()
        elif segments == 50:                                # strong scaling: seam rows to the neighbour, owned rows to rank 0
            h.ingest_device(t_rank.data_ptr(), t_off.data_ptr(), t_codes.data_ptr(), hi - lo)
            plan = gdist.seam_exchange_counts(h, int(d["rank"][hi - 1]) if hi > lo else -1)
            assert not plan.fallback and plan.bytes_on_wire() < 4 * (N + 2) * W * 49
        elif segments == -2:                                # packed exchange (uint16 lanes)
            h.ingest_device(t_rank.data_ptr(), t_off.data_ptr(), t_codes.data_ptr(), hi - lo)
            assert gdist.allreduce_counts_packed(h)
        elif segments == -3:                                # packed exchange behind the next job's expansion
            hb = Hansel.init_matrix(REF_SYMBOLS, REF_UNSYMBOLS, N, band_w=W, device=rank)
            hb.counts_buffer()

            def job(hh):
                hh.reset_counts()
                hh.ingest_device(t_rank.data_ptr(), t_off.data_ptr(), t_codes.data_ptr(), hi - lo)

            job(h)                                           # a representative job decides whether packing is safe
            assert gdist.OverlappedAllreduce.packing_is_safe(h)
            ov = gdist.OverlappedAllreduce([hb, h], free_sms=16, packed=True)
            for _ in range(4):                               # the jobs alternate between hb and h: the last one lands in h
                ov.step(job)
            ov.drain()
            # a count that cannot travel in a uint16 lane is flagged (the flag is summed with the data), not truncated
            with torch.cuda.stream(ov.main):
                job(hb)
                cptr, cn, _t, _tn = hb.counts_buffer()
                if rank == 1:
                    torch.as_tensor(gdist._DevBuf(cptr, cn, "<i4"), device=dev)[7] = 50000
                pptr, pn = hb.counts_pack(world)
                dist.all_reduce(torch.as_tensor(gdist._DevBuf(pptr, pn, "<i4"), device=dev), op=dist.ReduceOp.SUM)
            assert hb.counts_pack_overflowed()
            hb.close()
        elif segments > 1:
            pipe = gdist.PipelinedIngest(h, d["rank"][lo:hi], hi - lo, segments=segments)
            h.reset_counts()
            pipe.run(t_rank.data_ptr(), t_off.data_ptr(), t_codes.data_ptr())
        else:
            h.ingest_device(t_rank.data_ptr(), t_off.data_ptr(), t_codes.data_ptr(), hi - lo)
            gdist.allreduce_counts(h)
        s, c, v, sent = h.ingest_totals()
        h.finalize()
        util.set_totals(h, s, c, v)
        band = h.band()
        ref, rt = c_oracle.ingest(d["rank"], d["off"], d["codes"], N, W)
        assert (s, c, v, sent) == tuple(int(x) for x in rt)
        whole = segments != 50 or rank == 0                 # the seam exchange completes the matrix on rank 0 only
        if whole:
            assert np.array_equal(band, ref.astype(np.float32))
        # sharded public entry point from host arrays
        h2 = gdist.load_from_packed_sharded(d["rank"], d["off"], d["codes"], N, W, device=rank)
        assert np.array_equal(h2.band(), ref.astype(np.float32)) and h2.n_crumbs == c
        dist.barrier()
        open(out + str(rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("segments", [1, 4, 40, 50, -1, -2, -3])
def test_two_gpu_sharded_ingest(tmp_path, c_oracle, segments):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "ok")
    mp.spawn(_worker, args=(2, _free_port(), segments, out), nprocs=2, join=True)
    assert open(out + "0").read() == "ok" and open(out + "1").read() == "ok"
