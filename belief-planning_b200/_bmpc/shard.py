"""Episode sharding across ranks (one process per GPU).  Every episode is an independent problem with private persistent
state, so the solve path has NO collective: a rank owns a contiguous slice of the episode batch.  torch.distributed is
used only for the end-of-run statistics (count / sum / max reductions of a few scalars) and the timing barrier."""
import os


def world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_bounds(total, world_size, rank):
    """Contiguous slice [lo, hi) of `total` episodes owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(int(total), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_stats(stats, device=None):
    """All-reduce a dict of run statistics: keys starting with 'max_' use MAX, everything else SUM.  Works with the
    nccl backend (GPU tensors) and with gloo (CPU tensors); with a single process it returns the input."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(stats)
    keys = sorted(stats)
    out = {}
    for op, sel in ((dist.ReduceOp.MAX, [k for k in keys if k.startswith("max_")]),
                    (dist.ReduceOp.SUM, [k for k in keys if not k.startswith("max_")])):
        if not sel:
            continue
        t = torch.tensor([float(stats[k]) for k in sel], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=op)
        out.update({k: float(v) for k, v in zip(sel, t.tolist())})
    return out
