"""Batch sharding across ranks: frame pairs are independent units, so a batch is cut
into contiguous per-rank ranges and nothing is exchanged on the data path.  torch.distributed is
only the plumbing for barriers and for max/sum-over-ranks of scalars (NCCL on GPUs, gloo in tests)."""


def shard_range(npairs, rank, world):
    """Contiguous [first, last) of rank's pairs; sizes differ by at most one, earlier ranks larger."""
    if world < 1 or not (0 <= rank < world) or npairs < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(npairs, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def pair_seed(base_seed, global_pair_index):
    """Synthetic pair b of a job uses seed base + b, whatever rank holds it."""
    return base_seed + global_pair_index


def _reduce(x, op_name, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op_name))
    return float(t.item())


def max_over_ranks(x, device=None):
    return _reduce(x, "MAX", device)


def sum_over_ranks(x, device=None):
    return _reduce(x, "SUM", device)


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
