"""The N > 1 host path on CPU: world_size-2 gloo process group, weight-blob broadcast, frame/stream sharding, max-over-ranks
timing reduction (no GPU, no compute calls)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from desktop2stereo_b200 import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        blob = np.arange(10007, dtype=np.float32) * 0.5 if rank == 0 else None
        cfg = '{"hidden": 768}' if rank == 0 else None
        got, got_cfg = sharding.broadcast_weights(blob, cfg, src=0)
        frames = list(sharding.frames_for_rank(11, rank, world))
        streams = sharding.streams_for_rank(8, rank, world)
        slow = sharding.max_over_ranks(1.0 + rank)
        q.put((rank, float(got.sum()), got.dtype.str, got_cfg, frames, streams, slow))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = float((np.arange(10007, dtype=np.float32) * 0.5).sum())
    for rank, s, dt, cfg, frames, streams, slow in res:
        assert s == want and dt == "<f4" and cfg == '{"hidden": 768}'
        assert frames == list(range(rank, 11, 2))
        assert streams == [x for x in range(8) if x % 2 == rank]
        assert slow == 2.0                      # max over ranks
    all_frames = sorted(f for r in res for f in r[4])
    assert all_frames == list(range(11))        # every frame exactly once, no exchange needed


def test_single_process_is_identity():
    from desktop2stereo_b200 import sharding
    b = np.ones(5, np.float32)
    got, cfg = sharding.broadcast_weights(b, "{}")
    assert got is b and cfg == "{}"
    assert list(sharding.frames_for_rank(5, 0, 1)) == [0, 1, 2, 3, 4]
    assert sharding.max_over_ranks(3.5) == 3.5


def test_numa_binding_helpers():
    """bind_to_gpu_numa never raises (a box without CUDA / sysfs NUMA data stays unbound) and the cpulist parser handles ranges"""
    from desktop2stereo_b200.sharding import _parse_cpulist, bind_to_gpu_numa, streams_for_rank
    assert _parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11] and _parse_cpulist("") == []
    info = bind_to_gpu_numa(0)
    assert isinstance(info, dict) and "bound" in info
    # configs[4]: 8 streams over G ranks, every stream on exactly one rank
    for world in (1, 2, 4, 8):
        owned = [s for r in range(world) for s in streams_for_rank(8, r, world)]
        assert sorted(owned) == list(range(8)) and len(streams_for_rank(8, 0, world)) == 8 // world
