"""Multi-GPU plumbing: reference views are independent units of work
(app/smvsrecon.cc:658-733), so they are dealt round-robin to the ranks and
refined without any data-path collective. torch.distributed carries only
the bookkeeping (and, opt-in, the 272-double lighting reduction)."""
from __future__ import annotations

import numpy as np


def views_of_rank(n_views: int, rank: int, world: int):
    """View v goes to rank v mod world."""
    return list(range(rank, n_views, world))


def reduce_job_stats(pixel_iterations: float, seconds: float, device=None):
    """(sum over ranks of pixel-iterations, max over ranks of seconds)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(pixel_iterations), float(seconds)
    t = torch.tensor([pixel_iterations, seconds], dtype=torch.float64, device=device)
    s = t.clone()
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    m = t.clone()
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return float(s[0]), float(m[1])


def allreduce_lighting_normal_equations(A_b: np.ndarray, device=None) -> np.ndarray:
    """Opt-in global lighting: sum the 16x16 + 16 normal equations of
    LightOptimizer::fit_lighting_to_image over the ranks (the reference fits
    lighting per view and never shares it)."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.ascontiguousarray(A_b, dtype=np.float64), device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
