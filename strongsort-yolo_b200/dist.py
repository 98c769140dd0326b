"""Multi-GPU plumbing: the path shards by independent video stream (the reference's one
worker per --source, /root/reference/yolo_multi_model.py:351-354), one process per GPU,
NO data-path collective.  torch.distributed is used only for the barrier and for the
max-over-ranks of timed regions (NCCL on GPUs, gloo in the CPU tests) -- and for the one
optional exchange the north star names: the cross-stream ReID gallery (SharedGallery)."""
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


def gather_tracks(local_feat, local_ids):
    """all-gather every rank's exported tracks: [t_max, D] -> [G, t_max, D], [t_max] -> [G, t_max]
    (rank-major).  NCCL over NVLink / NVSwitch on GPUs (<= 513 KiB per rank at t_max 256), gloo on CPU."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_feat.unsqueeze(0), local_ids.unsqueeze(0)
    world = dist.get_world_size()
    all_feat = torch.empty((world,) + tuple(local_feat.shape), dtype=local_feat.dtype, device=local_feat.device)
    all_ids = torch.empty((world,) + tuple(local_ids.shape), dtype=local_ids.dtype, device=local_ids.device)
    # per-rank views of the stacked outputs: one ncclAllGather each on NCCL, and gloo accepts it too
    dist.all_gather(list(all_feat.unbind(0)), local_feat.contiguous())
    dist.all_gather(list(all_ids.unbind(0)), local_ids.contiguous())
    return all_feat, all_ids


def gather_packed(packed):
    """ONE collective for a stream's whole export: packed [L] (features + ids as 4-byte words) -> [G, L]."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return packed.unsqueeze(0)
    world = dist.get_world_size()
    flat = packed.contiguous().reshape(-1)
    out = torch.empty(world * flat.numel(), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, flat)                  # rank-major concatenation (NCCL and gloo)
    return out.view(world, flat.numel())


class SharedGallery:
    """Read-only cross-stream ReID gallery (BASELINE.json config C5): after a frame, ``step()`` exports
    this stream's confirmed tracks (csrc/gallery.cu), all-gathers every stream's export and matches
    each local track to the nearest track of another stream.  Nothing is written back into the
    tracker, so per-stream ids and the parity with the oracle are unaffected.

    Only the export (two tiny kernels reading the track table) runs on the tracker's stream; the
    all-gather and the match run on a side stream with double-buffered exports, so the lock-step the
    collective imposes on the ranks never stalls the per-stream association."""

    def __init__(self, tracker, t_max=256, max_dist=0.2, exchange="auto"):
        """exchange: "peer" -- the match kernel reads the other streams' exports straight out of their GPUs'
        memory over NVLink (torch symmetric memory: peer pointers + a device-side signal-pad barrier; no
        collective, no staging copy); "nccl" -- one all_gather_into_tensor of the packed export, then a local
        match; "auto" -- peer when symmetric memory can be set up for the group, else nccl."""
        import ctypes as C
        import torch
        from . import _lib
        self._C, self._torch, self._lib_mod = C, torch, _lib
        self.trk, self.t_max, self.max_dist = tracker, int(t_max), float(max_dist)
        dev, D = tracker.device, tracker.cfg.feat_dim
        z = lambda *shape, dt=torch.float32, fill=0: torch.full(shape, fill, dtype=dt, device=dev)
        # packed export: [t_max * D features | t_max ids | t_max slot scratch] as 4-byte words; the first
        # t_max * (D + 1) words are what the other streams read (513 KiB per rank per frame)
        self._symm = None
        self.exchange = "nccl"
        if exchange in ("auto", "peer"):
            try:
                self._symm = self._setup_peer(dev, D)
                self.exchange = "peer"
            except Exception as e:              # no symmetric memory here (single process, gloo, no P2P): NCCL path
                if exchange == "peer":
                    raise
                self._symm_error = repr(e)
        if self._symm is not None:
            self._packed = self._symm["bufs"]
        else:
            self._packed = [z(self.t_max * (D + 2), dt=torch.int32) for _ in range(2)]
        self._feat = [p[:self.t_max * D].view(torch.float32).view(self.t_max, D) for p in self._packed]
        self._ids2 = [p[self.t_max * D:] for p in self._packed]
        for i2 in self._ids2:
            i2.fill_(-1)
        self._count = [z(1, dt=torch.int32) for _ in range(2)]
        self._m_rank = [z(self.t_max, dt=torch.int32, fill=-1) for _ in range(2)]
        self._m_id = [z(self.t_max, dt=torch.int32, fill=-1) for _ in range(2)]
        self._m_dist = [z(self.t_max) for _ in range(2)]
        with torch.cuda.device(dev):
            self.stream = torch.cuda.Stream(device=dev)
            self._done = [torch.cuda.Event() for _ in range(2)]
        self._k = 0
        self._last = 0
        self.rank = env_rank_world()[0]

    def _setup_peer(self, dev, D):
        """Two symmetric-memory export buffers (double-buffered) + their rendezvous handles."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        if not (dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == "nccl"):
            raise RuntimeError("peer exchange needs an initialised NCCL group of more than one rank")
        words = self.t_max * (D + 2)
        bufs, hdls = [], []
        with torch.cuda.device(dev):
            for _ in range(2):
                b = symm_mem.empty(words, dtype=torch.int32, device=dev)
                b.zero_()
                hdls.append(symm_mem.rendezvous(b, dist.group.WORLD))
                bufs.append(b)
            best = torch.zeros(self.t_max, dtype=torch.int64, device=dev)
        torch.cuda.synchronize(dev)
        dist.barrier()
        return {"bufs": bufs, "hdls": hdls, "best": best}

    # views of the most recent step
    feat = property(lambda self: self._feat[self._last])
    ids = property(lambda self: self._ids2[self._last][:self.t_max])
    count = property(lambda self: self._count[self._last])
    m_rank = property(lambda self: self._m_rank[self._last])
    m_id = property(lambda self: self._m_id[self._last])
    m_dist = property(lambda self: self._m_dist[self._last])

    def step(self):
        """export (tracker stream) -> all-gather -> match (side stream); results stay on the device
        (``report()`` waits for and reads them)."""
        C, torch, _lib = self._C, self._torch, self._lib_mod
        trk, b = self.trk, self._k & 1
        lib = trk._lib                      # the library that owns the tracker handle
        feat, ids2, ids = self._feat[b], self._ids2[b], self._ids2[b][:self.t_max]
        with torch.cuda.device(trk.device):
            with torch.cuda.stream(trk.stream):
                trk.stream.wait_event(self._done[b])          # the side stream is done with this buffer
                # peer exchange: the other ranks read buffer b during THEIR match of two frames ago; they were past it
                # when last frame's barrier completed, which precedes last frame's match on our side stream
                trk.stream.wait_event(self._done[b ^ 1])
                _lib.check(lib.ssb_gallery_export(trk._h, self.t_max, _lib.ptr(feat), _lib.ptr(ids2),
                                                  _lib.ptr(self._count[b]), C.c_void_p(trk.stream.cuda_stream)),
                           "ssb_gallery_export")
            self.stream.wait_stream(trk.stream)
            with torch.cuda.stream(self.stream):
                D = int(feat.shape[1])
                if self._symm is not None:
                    # every rank's export of this frame is complete once all ranks passed the barrier (device-side,
                    # signal pads over NVLink); the kernel then pulls the foreign rows itself
                    hdl = self._symm["hdls"][b]
                    hdl.barrier(channel=0)
                    _lib.check(lib.ssb_gallery_peer_match(
                        C.c_void_p(int(hdl.buffer_ptrs_dev)), int(hdl.world_size), int(hdl.rank), self.t_max, D,
                        self.max_dist, _lib.ptr(self._symm["best"]), _lib.ptr(self._m_rank[b]), _lib.ptr(self._m_id[b]),
                        _lib.ptr(self._m_dist[b]), C.c_void_p(self.stream.cuda_stream)), "ssb_gallery_peer_match")
                else:
                    allp = gather_packed(self._packed[b][:self.t_max * (D + 1)])
                    G = int(allp.shape[0])
                    _lib.check(lib.ssb_gallery_cross_match_packed(
                        _lib.ptr(allp), G, self.rank if G > 1 else 0, self.t_max, D, self.max_dist,
                        _lib.ptr(self._m_rank[b]), _lib.ptr(self._m_id[b]), _lib.ptr(self._m_dist[b]),
                        C.c_void_p(self.stream.cuda_stream)), "ssb_gallery_cross_match_packed")
                    self._all = allp
                self._done[b].record(self.stream)
        self._last = b
        self._k += 1
        return self._m_rank[b], self._m_id[b], self._m_dist[b]

    def report(self):
        """[(local track id, remote rank, remote track id, cosine distance)] of the last step()."""
        self.stream.synchronize()
        n = int(self.count.item())
        ids, r, i, d = (x[:n].cpu().numpy() for x in (self.ids, self.m_rank, self.m_id, self.m_dist))
        return [(int(ids[k]), int(r[k]), int(i[k]), float(d[k])) for k in range(n) if r[k] >= 0]
