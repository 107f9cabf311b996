"""Batch sharding across ranks (one process per GPU; BASELINE.json configs[4]).

Proofs are independent, so the computation has no collective.  The wire records do move: rank 0 holds the batch
(n rln_witness_to_bytes_le records + n (r, s) pairs) in host memory, scatters contiguous slices to the ranks, every
rank proves its slice with the device-records entry point (rlnb200_prove_records_device: parsing, proving and
formatting all stay in HBM), and rank 0 gathers the fixed-size rln_proof_to_bytes_le records and brings them back to
host memory.  The collectives are torch.distributed's (NCCL over NVLink on the GPUs — ncclSend/ncclRecv groups under
dist.scatter / dist.gather; gloo in the CPU tests).

The in-process form of the same sharding (one process, one worker thread per device, no collective) is the C ABI's
rlnb200_multi_prove_batch.
"""
import torch
import torch.distributed as dist


def shard_bounds(total, world, rank):
    """contiguous slice [lo, hi) of `total` units owned by `rank`; the first total % world ranks get one extra"""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def scatter_records(full, record_bytes, total, device, src=0):
    """full: uint8 tensor of total*record_bytes on `src` (host or device; None elsewhere) → this rank's slice (uint8 tensor
    on `device`).  One host→device copy of the whole buffer on `src`; when the slices are equal the collective reads views
    of it, otherwise they are padded to a common size."""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes) * record_bytes
    out = torch.empty(width, dtype=torch.uint8, device=device)
    chunks = None
    if rank == src:
        d_full = full.to(device, non_blocking=True)
        if total % world == 0:
            chunks = list(d_full.view(world, width).unbind(0)) if width else [d_full[:0]] * world
        else:
            chunks = []
            for lo, hi in sizes:
                c = torch.zeros(width, dtype=torch.uint8, device=device)
                c[:(hi - lo) * record_bytes] = d_full[lo * record_bytes:hi * record_bytes]
                chunks.append(c)
    dist.scatter(out, chunks, src=src)
    lo, hi = sizes[rank]
    return out[:(hi - lo) * record_bytes]


def gather_records(local, record_bytes, total, dst=0):
    """inverse of scatter_records: the concatenated uint8 tensor of total*record_bytes on `dst` (same device as `local`),
    None elsewhere"""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes) * record_bytes
    if total % world == 0:
        padded = local
        whole = torch.empty(world * width, dtype=torch.uint8, device=local.device) if rank == dst else None
        bufs = list(whole.view(world, width).unbind(0)) if rank == dst and width else ([local[:0]] * world if rank == dst else None)
        dist.gather(padded.contiguous(), bufs, dst=dst)
        return whole if rank == dst else None
    padded = torch.zeros(width, dtype=torch.uint8, device=local.device)
    padded[:local.numel()] = local
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:(hi - lo) * record_bytes] for b, (lo, hi) in zip(bufs, sizes)])


def prove_sharded(prove_records, records, rs, total, rec_in, rec_out, device, out=None, src=0):
    """The whole multi-rank step.  On `src`: `records` (total × rec_in bytes) and `rs` (total × 64 bytes, or None for fresh
    randomness on every rank) are uint8 HOST tensors (pinned for speed); elsewhere they are ignored.

    prove_records(d_records, d_rs, n) -> uint8 device tensor of n × rec_out bytes (the rank-local prover:
    RLN.prove_records_device behind a tensor interface).

    Returns, on `src`, a uint8 host tensor of total × rec_out bytes (`out` if given, which should be pinned), None elsewhere."""
    rank = dist.get_rank()
    mine = scatter_records(records if rank == src else None, rec_in, total, device, src)
    has_rs = torch.tensor([1 if (rank == src and rs is not None) else 0], device=device)
    dist.broadcast(has_rs, src=src)
    my_rs = scatter_records(rs if rank == src else None, 64, total, device, src) if int(has_rs.item()) else None
    n = mine.numel() // rec_in
    proofs = prove_records(mine, my_rs, n)
    whole = gather_records(proofs.view(-1), rec_out, total, src)
    if rank != src:
        return None
    if out is None:
        out = torch.empty(total * rec_out, dtype=torch.uint8, pin_memory=whole.is_cuda)
    out.copy_(whole, non_blocking=False)
    return out


def rln_prove_records_fn(rln, rec_out):
    """adapter: an RLN handle → the prove_records callable of prove_sharded (device tensors in, device tensor out)"""
    def fn(d_records, d_rs, n):
        out = torch.empty(n * rec_out, dtype=torch.uint8, device=d_records.device)
        if n:
            st = torch.cuda.current_stream(d_records.device)
            rln.prove_records_device(d_records.data_ptr(), d_rs.data_ptr() if d_rs is not None else 0, n, out.data_ptr(), st.cuda_stream)
        return out
    return fn
