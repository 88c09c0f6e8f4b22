"""Multi-GPU plumbing: DP tasks (reads) are independent, so N GPUs = N processes,
each with its own context and a disjoint slice of the task stream; the only
cross-rank traffic is the reduction of the reported numbers (no data-path
collective; reference: the per-read independence of src/lamsa_aln.c:838-843)."""
import numpy as np


def shard_slices(n, world, chunk=4096):
    """Round-robin assignment of `chunk`-sized blocks of an n-task stream to ranks.
    Returns a list (per rank) of index arrays; blocks keep stream order inside a rank."""
    nblk = (n + chunk - 1) // chunk
    out = []
    for r in range(world):
        blocks = np.arange(r, nblk, world)
        idx = (blocks[:, None] * chunk + np.arange(chunk)[None, :]).reshape(-1)
        out.append(idx[idx < n])
    return out


def merge_results(parts, slices, n):
    """Inverse of shard_slices for per-task result arrays (structured or plain)."""
    first = parts[0]
    out = np.zeros(n, dtype=first.dtype)
    for p, s in zip(parts, slices):
        out[s] = p
    return out


def reduce_metrics(dist, device, time_like, sum_like):
    """MAX over ranks of the time-like values, SUM of the additive ones."""
    import torch
    t = torch.tensor(list(time_like), dtype=torch.float64, device=device)
    s = torch.tensor(list(sum_like), dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return t.tolist(), s.tolist()
