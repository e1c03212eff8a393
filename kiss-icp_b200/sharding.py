"""Multi-GPU layout of the hot path: independent sequences, one per GPU.

A KissICP instance is strictly sequential (scan k+1 needs the map and pose of scan k,
pipeline/KissICP.cpp:47,61-63) and its state is private (KissICP.hpp:87-95), so the only
parallel axis is ACROSS sequences: rank r of world W registers sequence r (weak scaling), with
no collective on the data path. The single exchange step is gathering the trajectories
(16 doubles per scan) at the end — torch.distributed all_gather (NCCL over NVLink on GPUs,
gloo in the CPU tests).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def sequences_of_rank(rank: int, world: int, n_sequences: int):
    """sequence ids owned by ``rank`` (round robin; n_sequences == world in the benchmark)"""
    return list(range(rank, n_sequences, world))


def gather_poses(local_poses: np.ndarray, device: str | torch.device = "cpu") -> np.ndarray:
    """(n_scans,4,4) per rank -> (world, n_scans, 4, 4) on every rank. Trajectories of all ranks
    have the same length in the benchmark (same --steps)."""
    t = torch.as_tensor(np.ascontiguousarray(local_poses), dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t.cpu().numpy()[None]
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()


def max_over_ranks(value: float, device: str | torch.device = "cpu") -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: str | torch.device = "cpu") -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
