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


def _parse_cpulist(text: str) -> list[int]:
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa(device_index: int, local_rank: int = 0, local_world: int = 1) -> dict:
    """Pin this process to the CPU cores of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is allocated, so that
    the per-frame staging buffers (first-touch) and the threads that fill them are local to the GPU's PCIe root.  Round 1 ran all
    8 ranks on NUMA node 0 (cores 0-31) and the 49.8 MB/frame device->host copies of 8 GPUs converged on one memory controller
    (SCALE_r01: e2e 0.24x at 8 GPUs).  When several ranks share a node its cores are split between them.
    Returns what was done (for the bench line); never raises — a container without sysfs/NUMA just stays unbound."""
    import os
    info = {"bound": False}
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        info.update(pci=bus, numa_node=node)
        if node < 0:
            return info
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return info
        # ranks whose GPUs share this node take disjoint slices of its cores
        peers = []
        for r in range(local_world):
            try:
                pr = torch.cuda.get_device_properties(r)
                b = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
                if int(open(f"/sys/bus/pci/devices/{b}/numa_node").read()) == node:
                    peers.append(r)
            except Exception:
                pass
        if local_rank in peers and len(peers) > 1 and len(allowed) >= len(peers):
            k = len(allowed) // len(peers)
            i = peers.index(local_rank)
            allowed = allowed[i * k:(i + 1) * k]
        os.sched_setaffinity(0, allowed)
        info.update(bound=True, cpus=len(allowed), first_cpu=allowed[0])
    except Exception as e:      # no sysfs, no NUMA, restricted cpuset ...
        info["error"] = f"{type(e).__name__}: {e}"
    return info
