"""Ensemble member management across GPUs: members shard over ranks in contiguous blocks, everything else is
replicated, and nothing is exchanged inside the time loop.  Collectives appear only where the reference's member loop has a
cross-member step: the final ensemble statistics here (all-reduce of sum / sum of squares / min / max of the hydrographs),
the EnKF analysis all-gather in enkf.py.  Works with any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in
the CPU tests); with no process group it degenerates to the single-rank result.
"""
import numpy as np


def shard_members(n_members, rank, world):
    """Contiguous block of global member indices owned by `rank`: (member0, n_local). Blocks differ by at most one member."""
    base, rem = divmod(int(n_members), int(world))
    n_local = base + (1 if rank < rem else 0)
    member0 = rank * base + min(rank, rem)
    return member0, n_local


def sample_shard(model, member0, n_local):
    """Replay the reference's member-serial parameter draws (CMonteCarloEnsemble::UpdateModel, one shared libc rand() stream)
    up to the end of this rank's block; returns the sampled values of the local members [n_local][n_dists] and leaves each
    local member's flattened parameter row refreshed (model.flatten(n_local) must have been called)."""
    model.ensemble_begin()
    out = []
    for e in range(member0 + n_local):
        v = model.ensemble_update_model(e)
        if e >= member0:
            model.flatten_member(e - member0)
            out.append(v)
    return np.array(out)


def ensemble_statistics(values, group=None):
    """Mean / standard deviation / min / max over ALL ensemble members of `values` [n_local_members, ...] (torch tensor on the
    device of the backend).  One all-reduce of [sum, sumsq, count] and one each for min and max."""
    import torch
    import torch.distributed as dist
    v = values.to(torch.float64)
    n = torch.tensor([float(v.shape[0])], dtype=torch.float64, device=v.device)
    s = v.sum(dim=0)
    mn = v.min(dim=0).values if v.shape[0] > 0 else torch.full_like(s, float("inf"))
    mx = v.max(dim=0).values if v.shape[0] > 0 else torch.full_like(s, float("-inf"))
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if multi:
        packed = torch.cat([s.reshape(-1), n])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        k = s.numel()
        s, n = packed[:k].reshape(s.shape), packed[k:]
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    mean = s / n
    d = v - mean
    m2 = (d * d).sum(dim=0)
    if multi:
        dist.all_reduce(m2, op=dist.ReduceOp.SUM, group=group)
    var = m2 / n
    return dict(mean=mean, std=torch.sqrt(var), min=mn, max=mx, n=int(n.item()))
