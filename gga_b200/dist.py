"""Frame sharding and scalar collectives (one process per GPU, torch.distributed).

The path shards by frame: every rank owns a contiguous block of frames and
runs the kernels on it with NO data-path collective.  The only exchanges are the ones the
reference itself does around the loss: a scalar all-reduce of the loss normaliser / log
scalars (mmdet ``reduce_mean`` in ``/root/reference/mmdet3d/models/dense_heads/
fcaf3d_head.py:298,306-308``; ``_parse_losses``) and, for the matching pass, the gather of
per-frame results to rank 0 (``tools/generate_pseudo_labels_gga.py:242``).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items, rank=None, world_size=None):
    """Contiguous block [lo, hi) of `n_items` frames owned by `rank` (sizes differ by <= 1)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_scalars(values, op='sum'):
    """All-reduces a small vector of scalars (<= 16 floats: loss sum, weight sum, n_pos ...)
    in ONE collective.  `values` is a 1-D tensor on the rank's device (CPU for gloo)."""
    rank, ws = world()
    if ws == 1:
        return values
    out = values.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM)
    if op == 'mean':
        out = out / ws
    return out


def reduce_mean(tensor):
    """mmdet ``reduce_mean``: all-reduce(sum) / world size (no-op without a process group)."""
    return reduce_scalars(tensor.reshape(-1), 'mean').reshape(tensor.shape)


def gather_frames(local, n_items_total):
    """All-gathers per-frame results (first dim = frames of this rank's shard) into the
    global frame order.  Shards may differ by one frame: padded to the max shard size."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_range(n_items_total, r, ws) for r in range(ws)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)
