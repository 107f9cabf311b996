"""CPU, world_size 2, gloo: the N>1 host logic (contiguous sharding, scatter of input records, gather of
proof records in order).  The per-rank "work" is a byte-wise stand-in; the GPU path itself is covered by -m gpu."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zerokit_b200.sharding import gather_records, prove_sharded, scatter_records, shard_bounds


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rec_in, rec_out = 1472, 288  # 46 input slots × 32 B in, proof + values out
    full = None
    if rank == 0:
        g = torch.Generator().manual_seed(1)
        full = torch.randint(0, 256, (total * rec_in,), dtype=torch.uint8, generator=g)
    mine = scatter_records(full, rec_in, total, torch.device("cpu"))
    lo, hi = shard_bounds(total, world, rank)
    assert mine.numel() == (hi - lo) * rec_in
    # stand-in for proving: first 288 bytes of each record, xor the global record index
    out = mine.view(hi - lo, rec_in)[:, :rec_out].clone()
    out[:, 0] ^= torch.arange(lo, hi, dtype=torch.int64).to(torch.uint8)
    res = gather_records(out.reshape(-1), rec_out, total)
    ok = True
    if rank == 0:
        want = full.view(total, rec_in)[:, :rec_out].clone()
        want[:, 0] ^= torch.arange(0, total, dtype=torch.int64).to(torch.uint8)
        ok = bool(torch.equal(res.view(total, rec_out), want))
    # the whole step as bench.py / a deployment runs it: host records on rank 0 → scatter → prove → gather → host records on
    # rank 0, with a stand-in prover that depends on the record AND on its (r, s) pair; uneven and even totals
    for tot in (total, 2 * world * 3):
        recs = rs = None
        if rank == 0:
            g = torch.Generator().manual_seed(tot)
            recs = torch.randint(0, 256, (tot * rec_in,), dtype=torch.uint8, generator=g)
            rs = torch.randint(0, 256, (tot * 64,), dtype=torch.uint8, generator=g)

        def fake_prover(d_records, d_rs, n):
            o = d_records.view(n, rec_in)[:, :rec_out].clone()
            o[:, 1] ^= d_rs.view(n, 64)[:, 5]
            return o.reshape(-1)
        got = prove_sharded(fake_prover, recs, rs, tot, rec_in, rec_out, torch.device("cpu"))
        if rank == 0:
            want = recs.view(tot, rec_in)[:, :rec_out].clone()
            want[:, 1] ^= rs.view(tot, 64)[:, 5]
            ok = ok and bool(torch.equal(got.view(tot, rec_out), want))
        else:
            ok = ok and got is None
    if rank == 0:
        q.put(ok)
    dist.destroy_process_group()


def test_scatter_prove_gather_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
