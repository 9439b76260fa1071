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


class PredictionGatherer(object):
    """The same all-gather for a stream of steps, off the critical path: equal shards [n_local, D] go into one of
    `depth` preallocated [world * n_local, D] buffers with an asynchronous all_gather_into_tensor, so the next
    step's forward does not wait for the collective (nor, through it, for the slowest rank of this step); a buffer
    is waited for before it is reused and by `result()` / `drain()`.  Works on NCCL (CUDA tensors; the collective
    runs on the backend's own stream, ordered after the producer of `local_out`) and on gloo (CPU tensors)."""

    def __init__(self, n_local, dim, dtype=torch.float32, device="cpu", depth=2, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.bufs = [torch.empty((self.world * int(n_local), int(dim)), dtype=dtype, device=device)
                     for _ in range(max(1, int(depth)))]
        self.work = [None] * len(self.bufs)
        self.n_local, self.i, self.last = int(n_local), 0, None

    def submit(self, local_out):
        """Start gathering this step's predictions; returns the slot they will land in."""
        if tuple(local_out.shape) != (self.n_local, self.bufs[0].shape[1]):
            raise ValueError("PredictionGatherer takes equal shards of shape %s" % ((self.n_local, self.bufs[0].shape[1]),))
        s = self.i % len(self.bufs)
        self.i += 1
        if self.work[s] is not None:
            self.work[s].wait()                     # the gather that last used this buffer
            self.work[s] = None
        if self.world == 1:
            self.bufs[s].copy_(local_out)
        else:
            self.work[s] = dist.all_gather_into_tensor(self.bufs[s], local_out.contiguous(), group=self.group,
                                                       async_op=True)
        self.last = s
        return s

    def result(self, slot=None):
        """The gathered [world * n_local, D] predictions of `slot` (default: the last submitted step), complete."""
        s = self.last if slot is None else slot
        if s is None:
            raise RuntimeError("PredictionGatherer.result before submit")
        if self.work[s] is not None:
            self.work[s].wait()
            self.work[s] = None
        return self.bufs[s]

    def drain(self):
        """Wait for every gather in flight (call before stopping a timer: the collectives are part of the job)."""
        for s, w in enumerate(self.work):
            if w is not None:
                w.wait()
                self.work[s] = None
