"""Batch sharding across GPUs: frames are independent in inference mode, so
the batch is split contiguously over ranks, parameters are replicated, and the only communication
is an all-gather of the [N_local, 69] predictions.  One process per GPU (torch.distributed)."""
import torch
import torch.distributed as dist


def shard_bounds(n_total, rank, world):
    """Contiguous [lo, hi) slice of the batch owned by `rank`; remainders go to the first ranks."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def gather_predictions(local_out, n_total, group=None):
    """all_gather of per-rank predictions [n_local, D] into [n_total, D] (rank order = batch order).
    Works on NCCL (CUDA tensors) and gloo (CPU tensors); ragged shards are padded to the largest."""
    if not (dist.is_available() and dist.is_initialized()):
        return local_out
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    n_max = max(hi - lo for lo, hi in sizes)
    buf = local_out
    if local_out.shape[0] < n_max:
        pad = torch.zeros((n_max - local_out.shape[0],) + tuple(local_out.shape[1:]),
                          dtype=local_out.dtype, device=local_out.device)
        buf = torch.cat([local_out, pad], 0)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf.contiguous(), group=group)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], 0)
