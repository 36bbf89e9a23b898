"""Multi-GPU plumbing of the hot path: one process per GPU, frames (or streams) sharded with NO per-frame collective.

The reference is single-device (no torch.distributed on the app path).  The only exchange this design needs
is getting the packed weight blob to every rank once at start-up: rank 0 packs it, one `broadcast` (NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests) delivers it.  After that every rank owns a whole engine and a disjoint
subset of the work:
  * independent frames (EMA off)      -> frame i goes to rank i % world            (`frames_for_rank`)
  * video streams (EMA / VDA state)   -> stream s goes to rank s % world, its frames stay in order (`streams_for_rank`)
"""
from __future__ import annotations

import json

import numpy as np
import torch
import torch.distributed as dist


def frames_for_rank(n_frames: int, rank: int, world: int) -> range:
    """Round-robin frame sharding: frame i -> rank i % world."""
    return range(rank, n_frames, world)


def streams_for_rank(n_streams: int, rank: int, world: int) -> list[int]:
    """Stateful streams (DepthStabilizer EMA, depth.py:1865-1887) stay on one rank: stream s -> rank s % world."""
    return [s for s in range(n_streams) if s % world == rank]


def broadcast_weights(blob: np.ndarray | None, config_json: str | None, src: int = 0, device=None):
    """Rank `src` passes the packed fp32 blob (weights.pack_state_dict) and the HF config JSON; every rank returns
    (blob ndarray, config_json).  One collective at init, none per frame."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return blob, config_json
    rank = dist.get_rank()
    meta = [config_json, int(blob.size) if rank == src else None]
    dist.broadcast_object_list(meta, src=src)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    if rank == src:
        t = torch.from_numpy(np.ascontiguousarray(blob, dtype=np.float32)).to(dev)
    else:
        t = torch.empty(meta[1], dtype=torch.float32, device=dev)
    dist.broadcast(t, src=src)
    return t.cpu().numpy(), meta[0]


def max_over_ranks(value: float, device=None) -> float:
    """Timing reduction used by bench.py: the job is as slow as its slowest rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
