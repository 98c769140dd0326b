"""Multi-GPU plumbing: the path shards by independent video stream (the reference's one
worker per --source, /root/reference/yolo_multi_model.py:351-354), one process per GPU,
NO data-path collective.  torch.distributed is used only for the barrier and for the
max-over-ranks of timed regions (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def streams_for_rank(rank, world, n_streams):
    """stream i -> rank i mod world (C3: 4 streams / 4 GPUs, C5: 8 / 8)."""
    return [s for s in range(n_streams) if s % world == rank]


def init(backend, device=None):
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    if not dist.is_initialized():
        kw = {"device_id": device} if (device is not None and backend == "nccl") else {}
        dist.init_process_group(backend, **kw)
    return dist


def max_over_ranks(x, device="cpu"):
    """max of a python float over all ranks (timed regions are reported as the slowest rank)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_fps(frames_per_rank, seconds_max, world):
    """whole-job throughput: all ranks' frames over the slowest rank's time."""
    return world * frames_per_rank / seconds_max
