"""Batch sharding across ranks (one process per GPU).  Proofs are independent, so the data path has no
collective: rank 0 scatters the per-proof input records, every rank proves its contiguous slice, rank 0
gathers the fixed-size proof records.  Works with any torch.distributed backend (NCCL on the
GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_bounds(total, world, rank):
    """contiguous slice [lo, hi) of `total` units owned by `rank`; the first total % world ranks get one extra"""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def scatter_records(full, record_bytes, total, device, src=0):
    """full: uint8 tensor of total*record_bytes on `src` (None elsewhere) → this rank's slice (uint8 tensor on device).
    Slices may differ in length by one record; they are padded to a common size for the collective."""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes) * record_bytes
    out = torch.empty(width, dtype=torch.uint8, device=device)
    chunks = None
    if rank == src:
        chunks = []
        for lo, hi in sizes:
            c = torch.zeros(width, dtype=torch.uint8, device=device)
            c[:(hi - lo) * record_bytes] = full[lo * record_bytes:hi * record_bytes].to(device)
            chunks.append(c)
    dist.scatter(out, chunks, src=src)
    lo, hi = sizes[rank]
    return out[:(hi - lo) * record_bytes]


def gather_records(local, record_bytes, total, dst=0):
    """inverse of scatter_records: returns the concatenated uint8 tensor of total*record_bytes on `dst`, None elsewhere"""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes) * record_bytes
    padded = torch.zeros(width, dtype=torch.uint8, device=local.device)
    padded[:local.numel()] = local
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:(hi - lo) * record_bytes] for b, (lo, hi) in zip(bufs, sizes)])
