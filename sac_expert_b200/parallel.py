"""Multi-GPU plumbing (one process per GPU, torch.distributed).

(a) Population sharding - the natural parallelism of this path: agents are independent (own
    parameters, replay, expert rows), mirroring the reference's one-process-per-seed pool
    (``/root/reference/sac_eo/train.py:130-152``).  Agent ``i`` lives on rank ``i % world``; there is NO
    per-step collective.
(b) Optional single-agent data-parallel mode: the SAME agents on every rank, the minibatch rows split
    across ranks; gradients are averaged with an all-reduce between the phase-split kernels
    (``saceo_update_phase``): critics -> actor -> alpha, three dependent collectives per update.
    Every rank normalises its losses by its LOCAL row count, so the average over equally sized
    slices equals the global-batch mean the reference computes (``reduce_mean`` over B,
    ``SAC_expert.py:241,319,345``).  The expert term is replicated on every rank (same E rows, same
    draws), so averaging leaves it unchanged.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def agent_shard(n_total: int, rank: int, world: int) -> List[int]:
    """Global agent ids owned by ``rank`` (round-robin, like ``i % n_gpu`` in SURVEY.md §8e)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_total, world))


def shard_sizes(n_total: int, world: int) -> List[int]:
    return [len(agent_shard(n_total, r, world)) for r in range(world)]


def slice_rows(x: np.ndarray, rank: int, world: int, axis: int = 0) -> np.ndarray:
    """Rank slice of ONE global draw (indices / noise rows), so that the union over ranks is exactly
    the single-GPU minibatch."""
    n = x.shape[axis]
    if n % world:
        raise ValueError(f"batch of {n} rows does not split evenly over {world} ranks")
    per = n // world
    sl = [slice(None)] * x.ndim
    sl[axis] = slice(rank * per, (rank + 1) * per)
    return x[tuple(sl)]


def dp_noise_for_rank(noise: np.ndarray, B: int, E: int, rank: int, world: int) -> np.ndarray:
    """Per-agent noise block [3B+E, A] (u1 | u2 | uE | u5) -> the rank's block [3B/w+E, A]: the three
    B-row draws are row-sliced, the expert draw is replicated."""
    u1, u2, uE, u5 = noise[:B], noise[B:2 * B], noise[2 * B:2 * B + E], noise[2 * B + E:]
    return np.concatenate([slice_rows(u1, rank, world), slice_rows(u2, rank, world), uE,
                           slice_rows(u5, rank, world)], 0)


def average_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean over the ranks of ``group`` (NCCL on GPU tensors, gloo on CPU tensors)."""
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.div_(dist.get_world_size(group))
    return t


def dp_update(pop, num_timesteps: int = 0, group=None) -> torch.Tensor:
    """One data-parallel update of ``pop`` (whose ``B`` is the LOCAL slice size).  Draws must have been
    injected with ``pop.set_draws`` (rank-sliced from one global draw)."""
    g_q = pop.debug("g_q")
    g_a = pop.debug("g_actor")
    losses = pop.debug("losses")
    pop.update_phase(0, num_timesteps)        # TD target + critic grads (local rows)
    average_(g_q, group)
    pop.update_phase(1, num_timesteps)        # critic Adam + Polyak (identical on every rank)
    pop.update_phase(2, num_timesteps)        # actor grads through the updated critics
    average_(g_a, group)
    pop.update_phase(3, num_timesteps)        # actor Adam
    pop.update_phase(4, num_timesteps)        # alpha gradient (last word of g_actor)
    last = g_a.view(pop.spec.n_agents, -1)[:, -1].clone()
    average_(last, group)
    g_a.view(pop.spec.n_agents, -1)[:, -1] = last
    pop.update_phase(5, num_timesteps)        # alpha Adam + clamp
    average_(losses, group)
    return losses.view(pop.spec.n_agents, -1)
