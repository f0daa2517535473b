"""Host-side sharding of independent code blocks over ranks (one process per GPU).

The path has no exchange step: code blocks are partitioned contiguously, each rank decodes its own
range on its own GPU, and only results (decoded bytes, CRC flags, iteration counts) and timings are
gathered over torch.distributed -- no data-path collective.
"""
import numpy as np


def shard_bounds(n_items, world, rank):
    """[first, last) of the contiguous range rank `rank` of `world` owns: block i -> rank floor(i*world/n)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    first = (n_items * rank) // world
    last = (n_items * (rank + 1)) // world
    return first, last


def balanced_bounds(costs, world):
    """Contiguous partition of per-block costs (e.g. K * expected iterations) into `world` ranges with
    near-equal total cost; returns world+1 boundaries."""
    c = np.concatenate([[0], np.cumsum(np.asarray(costs, dtype=np.float64))])
    total = c[-1]
    b = [0]
    for r in range(1, world):
        b.append(int(np.searchsorted(c, total * r / world, side="left")))
    b.append(len(costs))
    return [min(max(x, 0), len(costs)) for x in b]


def gather_to_rank0(dist, local_np, counts, device=None):
    """Gather per-rank numpy arrays (first axis = this rank's blocks) to rank 0, concatenated in rank order.
    Works with the gloo (CPU tensors) and nccl (device tensors) backends.  Returns None on other ranks."""
    import torch
    world = dist.get_world_size()
    rank = dist.get_rank()
    mx = int(max(counts))
    pad_shape = (mx,) + tuple(local_np.shape[1:])
    buf = np.zeros(pad_shape, local_np.dtype)
    buf[: local_np.shape[0]] = local_np
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    if rank != 0:
        return None
    return np.concatenate([o.cpu().numpy()[: counts[r]] for r, o in enumerate(outs)], axis=0)


def max_over_ranks(dist, value, device=None):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
