"""``StrongSORT.update(dets, img)`` -- host-side mirror of the reference's tracker
seam, driving the sm_100a kernels of libssb.so through the C-ABI (include/ssb.h).

Reference boundary: the tracker call of /root/reference/yolo_multi_model.py:41
(``model.track(..., persist=True, tracker=...)``) behind which upstream's
``StrongSORT.update(dets, ori_img)`` sits (the strong_sort/
package itself is absent from the snapshot).  Same constructor knobs
(strong_sort.yaml, A.1), same argument meaning, same output rows
``[x1, y1, x2, y2, track_id, class_id, conf]``.

PyTorch is used for device memory, pinned staging buffers and streams only.
There is no CPU path: without libssb.so or a CUDA device construction raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, weights as _weights

CNT_OUT_ROWS, CNT_TRACKS, CNT_CONFIRMED, CNT_NEXT_ID = 0, 1, 2, 3
CNT_MATCHES_A, CNT_MATCHES_B, CNT_NEW, CNT_ERROR = 4, 5, 6, 7
_HDR_BYTES = 64          # counts live in front of the output rows


class StrongSORT:
    def __init__(self, model_weights=None, device="cuda:0", fp16=False,
                 max_dist=0.2, max_iou_distance=0.7, max_age=30, n_init=3,
                 nn_budget=100, mc_lambda=0.995, ema_alpha=0.9,
                 max_tracks=1024, max_dets=512, reid_backend="tc", debug=False):
        """``reid_backend`` other than "tc" and ``debug=True`` (``reid_block``, ``debug_costs``, phase stamps) load
        libssb_dbg.so -- the product library libssb.so carries neither the baselines nor the diagnostics."""
        torch = _lib.require_cuda()
        self._torch = torch
        self._debug = bool(debug) or reid_backend != "tc"
        self._lib = _lib.load(debug=self._debug)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.SsbError("StrongSORT(device=...) must be a CUDA device")
        self.fp16 = bool(fp16)   # accepted for signature parity; kernels pick their own precision
        cfg = _lib.SsbConfig()
        self._lib.ssb_default_config(C.byref(cfg))
        cfg.max_tracks, cfg.max_dets = int(max_tracks), int(max_dets)
        cfg.nn_budget, cfg.n_init, cfg.max_age = int(nn_budget), int(n_init), int(max_age)
        cfg.max_dist, cfg.max_iou_distance = float(max_dist), float(max_iou_distance)
        cfg.mc_lambda, cfg.ema_alpha = float(mc_lambda), float(ema_alpha)
        self.cfg = cfg
        nbytes = self._lib.ssb_workspace_bytes(C.byref(cfg))
        if nbytes < 0:
            _lib.check(-1, "ssb_workspace_bytes")
        with torch.cuda.device(self.device):
            # high priority: when the association of frame k shares the GPU with the OSNet of frame k+1 (two-stage
            # pipeline), its many small kernels get the SM slots the big ReID CTAs free up first
            self.stream = torch.cuda.Stream(device=self.device, priority=-1)
            self._ws = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=self.device)
            base = (self._ws.data_ptr() + 255) & ~255
            h = C.c_void_p()
            _lib.check(self._lib.ssb_create(C.byref(cfg), C.c_void_p(base), nbytes, C.byref(h)),
                       "ssb_create")
            self._h = h
            # ReID weights: fp16 hi/lo operand blob of the tensor-core kernels (+ the fp32 blob of the
            # SIMT baseline when the debug library is loaded)
            sd = _weights.load_state_dict(model_weights)
            if self._debug:
                blob, sizes = _weights.pack(_weights.fold(sd))
                self._w_blob = torch.from_numpy(blob).to(self.device)
                self._w_sizes = (C.c_int64 * len(sizes))(*[int(s) for s in sizes])
                _lib.check(self._lib.ssb_reid_set_weights(self._h, _lib.ptr(self._w_blob),
                                                          self._w_sizes, len(sizes)),
                           "ssb_reid_set_weights")
            tc_blob, tc_off = _weights.pack_tc(_weights.fold(sd))
            self._w_tc = torch.from_numpy(tc_blob).to(self.device)
            self._w_tc_off = (C.c_int64 * len(tc_off))(*[int(o) for o in tc_off])
            _lib.check(self._lib.ssb_reid_set_weights_tc(self._h, _lib.ptr(self._w_tc),
                                                         self._w_tc_off, len(tc_off)), "ssb_reid_set_weights_tc")
            self.set_reid_backend(reid_backend)
            # staging
            S, N = cfg.max_tracks, cfg.max_dets
            self._dets_pin = torch.empty((N, 6), dtype=torch.float32).pin_memory()
            self._dets_dev = torch.empty((N, 6), dtype=torch.float32, device=self.device)
            self._out_bytes = _HDR_BYTES + S * _lib.SSB_OUT_COLS * 8
            self._out_dev = torch.zeros(self._out_bytes, dtype=torch.uint8, device=self.device)
            self._out_pin = torch.zeros(self._out_bytes, dtype=torch.uint8).pin_memory()
            self._img_dev = None
            self._img_pin = None
            self._feats_dev = torch.empty((N, cfg.feat_dim), dtype=torch.float32, device=self.device)
        self._out_np = self._out_pin.numpy()
        self._counts_np = self._out_np[:32].view(np.int32)
        self._rows_np = self._out_np[_HDR_BYTES:].view(np.float64).reshape(S, _lib.SSB_OUT_COLS)
        self._track_hint = 0
        self.last_counts = np.zeros(8, dtype=np.int32)
        self.last_det_index = np.zeros(0, dtype=np.int64)
        self.reset()

    # ------------------------------------------------------------------
    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.ssb_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_reid_backend(self, name):
        """'tc' (default): tcgen05 OSBlocks on fp16 hi/lo operand planes with DSMEM halo exchange
        (csrc/reid_tc4.cu); 'tc3': round-1 pointwise GEMM + fp32 depthwise kernel on float32 NHWC
        activations with recomputed halos (csrc/reid_tc3.cu); 'tc9': LightConv as 9 shifted tcgen05
        GEMMs (csrc/reid_tc.cu); 'simt': fp32 CUDA-core baseline.  The last three are A/B baselines."""
        modes = {"simt": 0, "tc9": 1, "tc3": 2, "tc": 3}
        if name not in modes:
            raise ValueError("reid_backend must be 'tc', 'tc3', 'tc9' or 'simt'")
        if not self._debug:
            if name != "tc":
                raise _lib.SsbError("the baselines live in libssb_dbg.so: construct StrongSORT(debug=True)")
        else:
            _lib.check(self._lib.ssb_reid_use_tc(self._h, modes[name]), "ssb_reid_use_tc")
        self.reid_backend = name

    def _need_debug(self, what):
        if not self._debug:
            raise _lib.SsbError(f"{what} needs libssb_dbg.so: construct StrongSORT(debug=True)")

    def reid_tc_status(self):
        """0 when no tensor-core barrier wait ever timed out."""
        v = C.c_int32(0)
        _lib.check(self._lib.ssb_reid_tc_status(self._h, C.byref(v), C.c_void_p(self.stream.cuda_stream)),
                   "ssb_reid_tc_status")
        return int(v.value)

    def reid_block(self, block, x, use_tc):
        """One OSBlock on a float32 NHWC array [n,H,W,cin] (parity tests); use_tc: False/0 simt,
        1 'tc9' kernel, 2 'tc3' kernel, True/3 'tc' kernel (operand planes; converted in and out)."""
        use_tc = 3 if use_tc is True else int(use_tc)
        self._need_debug("reid_block")
        torch = self._torch
        couts = [64, 64, 96, 96, 128, 128]
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).to(self.device)
            n, H, W, _ = xd.shape
            yd = torch.zeros((n, H, W, couts[block]), dtype=torch.float32, device=self.device)
            _lib.check(self._lib.ssb_reid_block(self._h, int(block), _lib.ptr(xd), _lib.ptr(yd), int(n),
                                                use_tc, C.c_void_p(self.stream.cuda_stream)),
                       "ssb_reid_block")
        self.stream.synchronize()
        return yd.cpu().numpy()

    def reset(self):
        """Forget all tracks (ids restart at 1)."""
        _lib.check(self._lib.ssb_reset(self._h, C.c_void_p(self.stream.cuda_stream)), "ssb_reset")
        self.stream.synchronize()
        self._track_hint = 0

    # ------------------------------------------------------------------
    def _stage_image(self, img, caller=None):
        torch = self._torch
        if torch.is_tensor(img):
            if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3:
                raise ValueError("img tensor must be uint8 [H,W,3]")
            if img.is_cuda:
                if img.device != self.device:
                    img = img.to(self.device)
                if caller is not None:
                    self._after_producer(img, self.stream, caller)
                return img.contiguous()
            src = img.contiguous()
        else:
            img = np.asarray(img)
            if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
                raise ValueError("img must be a uint8 HxWx3 array (BGR)")
            src = torch.from_numpy(np.ascontiguousarray(img))
        if self._img_dev is None or tuple(self._img_dev.shape) != tuple(src.shape):
            self._img_dev = torch.empty(tuple(src.shape), dtype=torch.uint8, device=self.device)
        self._img_dev.copy_(src, non_blocking=True)      # on self.stream (set by caller)
        return self._img_dev

    def prefetch(self, ori_img):
        """Start the host->device copy of a frame NOW (asynchronously, on the tracker's stream) so that it overlaps
        whatever the caller does before ``update(dets, ori_img)`` -- typically the detector post-process that
        produces ``dets``.  ``update`` recognises the same frame object and skips its own copy."""
        torch = self._torch
        if torch.is_tensor(ori_img) and ori_img.is_cuda:
            return
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self._staged = (ori_img, self._stage_image(ori_img))

    def _after_producer(self, t, stream, caller):
        """A CUDA tensor handed in by the caller may still be pending on the caller's stream (e.g.
        YoloNMS output): order our stream after it and keep the allocator from recycling it.
        ``caller`` is the stream that was current BEFORE our own stream context was entered."""
        stream.wait_stream(caller)
        t.record_stream(stream)

    @staticmethod
    def _raise_on_error(cnt):
        err = int(cnt[CNT_ERROR])
        if err & 2:
            raise _lib.SsbError("a tensor-core barrier wait timed out inside the ReID kernels "
                                "(device status word set): this frame's embeddings are invalid")
        if err & 1:
            raise _lib.SsbError("track table overflow: raise max_tracks / max_dets")

    def increment_ages(self):
        """Upstream ``StrongSORT.increment_ages()``: what its stream loop calls instead of ``update``
        on a frame without detections (ages advance, tracks are marked missed, no Kalman predict)."""
        _lib.check(self._lib.ssb_increment_ages(self._h, C.c_void_p(self.stream.cuda_stream)),
                   "ssb_increment_ages")

    def class_counts(self):
        """The reference's ``--count`` overlay numbers (yolo_multi_model.py:284-300) from the device track
        table: {class id: number of track ids whose reported rows carry that class most often}."""
        torch = self._torch
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            out = torch.zeros(80, dtype=torch.int32, device=self.device)
            _lib.check(self._lib.ssb_class_counts(self._h, _lib.ptr(out), C.c_void_p(self.stream.cuda_stream)),
                       "ssb_class_counts")
            host = out.cpu()
        self.stream.synchronize()
        return {int(c): int(v) for c, v in enumerate(host.tolist()) if v}

    def update(self, dets, ori_img, features=None):
        """dets: [N,6] (x1,y1,x2,y2,conf,cls) torch CPU tensor / ndarray (or a
        CUDA tensor); ori_img: BGR uint8 HxWx3 ndarray (or torch tensor, CPU
        pinned or CUDA).  Returns float64 ndarray [M,7]; ``[]``-like empty
        array when there is nothing to report (caller treats as "no ids",
        yolo_multi_model.py:54).  ``features`` (tests only) bypasses OSNet."""
        torch = self._torch
        lib = self._lib
        if torch.is_tensor(dets):
            d = dets.detach()
        else:
            d = torch.from_numpy(np.ascontiguousarray(np.asarray(dets, dtype=np.float32)))
        d = d.reshape(-1, 6).to(torch.float32)
        n = int(d.shape[0])
        if n > self.cfg.max_dets:
            raise ValueError(f"{n} detections exceed max_dets={self.cfg.max_dets}")
        H, W = int(ori_img.shape[0]), int(ori_img.shape[1])
        st = C.c_void_p(self.stream.cuda_stream)
        caller = torch.cuda.current_stream(self.device)      # the stream the caller's tensors were produced on
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            if n:
                if d.is_cuda:
                    self._after_producer(d, self.stream, caller)
                    self._dets_dev[:n].copy_(d, non_blocking=True)
                else:
                    self._dets_pin[:n].copy_(d)
                    self._dets_dev[:n].copy_(self._dets_pin[:n], non_blocking=True)
            feats_ptr = None
            img_dev = None
            if features is not None:
                f = torch.as_tensor(features, dtype=torch.float32).reshape(n, self.cfg.feat_dim)
                if f.is_cuda:
                    self._after_producer(f, self.stream, caller)
                self._feats_dev[:n].copy_(f, non_blocking=False)
                feats_ptr = _lib.ptr(self._feats_dev)
            elif n:
                staged = getattr(self, "_staged", None)
                if staged is not None and staged[0] is ori_img:
                    img_dev = staged[1]                      # prefetch() already enqueued the copy on this stream
                else:
                    img_dev = self._stage_image(ori_img, caller)
            self._staged = None
            pitch = 3 * W
            _lib.check(lib.ssb_update(
                self._h, _lib.ptr(self._dets_dev), n,
                _lib.ptr(img_dev) if img_dev is not None else None, H, W, pitch,
                feats_ptr, C.c_void_p(self._out_dev.data_ptr() + _HDR_BYTES),
                _lib.ptr(self._out_dev), self._track_hint, st), "ssb_update")
            self._out_pin.copy_(self._out_dev, non_blocking=True)
        self.stream.synchronize()
        cnt = self._counts_np
        self.last_counts = cnt.copy()
        self._raise_on_error(cnt)
        self._track_hint = int(cnt[CNT_TRACKS])
        m = int(cnt[CNT_OUT_ROWS])
        rows = self._rows_np[:m].copy()
        self.last_det_index = rows[:, 7].astype(np.int64)
        return rows[:, :7]

    # ------------------------------------------------------------------
    def camera_update(self, previous_img, current_img, warp_matrix=None):
        """Upstream ``tracker.camera_update(prev, curr)``, called by the stream loop
        between frames when ECC is on.  Upstream runs cv2.findTransformECC once PER TRACK on the
        same image pair; here the 2x3 warp is estimated once per frame (``ecc.ecc_warp``: same
        0.1x grayscale, MOTION_EUCLIDEAN, eps 1e-5, 100 iterations) -- or passed in as
        ``warp_matrix`` -- and applied to every live track by one kernel."""
        if warp_matrix is None:
            from . import ecc
            warp_matrix = ecc.ecc_warp(previous_img, current_img)
            if warp_matrix is None:
                return None
        w = np.ascontiguousarray(np.asarray(warp_matrix, dtype=np.float64).reshape(2, 3))
        arr = (C.c_double * 6)(*w.reshape(-1).tolist())
        _lib.check(self._lib.ssb_camera_update(self._h, arr, C.c_void_p(self.stream.cuda_stream)),
                   "ssb_camera_update")
        return w

    # ------------------------------------------------------------------
    def extract_features(self, ori_img, boxes_xyxy_int):
        """OSNet embeddings of integer crop boxes (parity tests / ssb_reid)."""
        torch = self._torch
        b = torch.as_tensor(np.asarray(boxes_xyxy_int, dtype=np.int32).reshape(-1, 4))
        n = int(b.shape[0])
        H, W = int(ori_img.shape[0]), int(ori_img.shape[1])
        caller = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            img_dev = self._stage_image(ori_img, caller)
            bd = b.to(self.device)
            out = torch.empty((n, self.cfg.feat_dim), dtype=torch.float32, device=self.device)
            _lib.check(self._lib.ssb_reid(self._h, _lib.ptr(img_dev), H, W, 3 * W, _lib.ptr(bd), n,
                                          _lib.ptr(out), C.c_void_p(self.stream.cuda_stream)),
                       "ssb_reid")
        self.stream.synchronize()
        return out.cpu().numpy()

    # ------------------------------------------------------------------
    # two-stage pipeline: embedding of frame k overlaps the association of frame k-1
    # ------------------------------------------------------------------
    def _pipe_init(self, lag=1):
        torch = self._torch
        if lag not in (1, 2, 3):
            raise ValueError("lag must be 1, 2 or 3 frames")
        R = lag + 1                                  # result buffers: frames k-lag .. k are in flight
        with torch.cuda.device(self.device):
            # one embedding stream per slot: the OSNet launches of two consecutive frames may
            # interleave on the GPU and fill each other's partial waves
            self._pstreams = [torch.cuda.Stream(device=self.device) for _ in range(2)]
            self._pstream = self._pstreams[0]
            N = self.cfg.max_dets
            self._p_dets = [torch.empty((N, 6), dtype=torch.float32, device=self.device) for _ in range(2)]
            self._p_dets_pin = [torch.empty((N, 6), dtype=torch.float32).pin_memory() for _ in range(2)]
            # host frames: their H2D copies run on a stream of their own into a ring of three device buffers, so the
            # copy of frame k+1 is under way while frame k is embedded (a buffer is free again once the embedding
            # that read it is done -- not only once that frame has been associated)
            self._p_copy = torch.cuda.Stream(device=self.device)
            self._p_img = [None, None, None]
            self._p_img_ready = [torch.cuda.Event() for _ in range(3)]
            self._p_img_free = [torch.cuda.Event() for _ in range(3)]
            self._p_out = [torch.zeros(self._out_bytes, dtype=torch.uint8, device=self.device) for _ in range(R)]
            self._p_pin = [torch.zeros(self._out_bytes, dtype=torch.uint8).pin_memory() for _ in range(R)]
            self._p_embed_done = [torch.cuda.Event() for _ in range(2)]
            self._p_assoc_done = [torch.cuda.Event() for _ in range(2)]     # per embedding slot (GPU-side slot reuse)
            self._p_res_done = [torch.cuda.Event() for _ in range(R)]       # per result buffer (host-side collect)
        self._p_lag, self._p_ring = lag, R
        self._p_k = 0
        self._p_done = 0                             # frames collected so far
        self._p_n = {}                               # detections of the frames not yet collected

    def _pipe_collect(self, k):
        """Wait for the association of frame k of the pipeline; return its rows."""
        r = k % self._p_ring
        self._p_res_done[r].synchronize()
        buf = self._p_pin[r].numpy()
        cnt = buf[:32].view(np.int32)
        self.last_counts = cnt.copy()
        self._p_done = k + 1
        self._p_n.pop(k, None)
        self._raise_on_error(cnt)
        self._track_hint = int(cnt[CNT_TRACKS])
        m = int(cnt[CNT_OUT_ROWS])
        rows = buf[_HDR_BYTES:].view(np.float64).reshape(-1, _lib.SSB_OUT_COLS)[:m].copy()
        self.last_det_index = rows[:, 7].astype(np.int64)
        return rows[:, :7]

    def update_pipelined(self, dets, ori_img, lag=1):
        """Same arguments as ``update``; returns the rows of the frame submitted ``lag`` calls ago (None while
        fewer have been submitted).  ``lag=1`` (default): one frame of latency buys the overlap of this frame's
        OSNet with the previous frame's association on a second stream.  ``lag=2`` additionally takes the host out
        of the loop: with one frame of latency the caller cannot hand over frame k+1 before frame k-1's rows came
        back, so every second frame's OSNet starts a host round trip late; with two, every dependency between frames
        is a device-side event.  ``lag`` is fixed by the first call (``flush_pipelined`` / ``drain_pipelined`` reset
        it).  Results are identical to ``update``."""
        torch = self._torch
        if not hasattr(self, "_pstream") or (self._p_k == 0 and self._p_lag != lag):
            self._pipe_init(lag)
        elif lag != self._p_lag:
            raise ValueError("lag changes only on an empty pipeline (call drain_pipelined() first)")
        k, slot = self._p_k, self._p_k & 1
        if torch.is_tensor(dets):
            d = dets.detach().reshape(-1, 6).to(torch.float32)
        else:
            d = torch.from_numpy(np.ascontiguousarray(np.asarray(dets, dtype=np.float32))).reshape(-1, 6)
        n = int(d.shape[0])
        if n > self.cfg.max_dets:
            raise ValueError(f"{n} detections exceed max_dets={self.cfg.max_dets}")
        H, W = int(ori_img.shape[0]), int(ori_img.shape[1])
        pst = self._pstreams[slot]
        caller = torch.cuda.current_stream(self.device)
        host_img = n > 0 and not (torch.is_tensor(ori_img) and ori_img.is_cuda)
        ri = k % 3
        if host_img:
            src = ori_img if torch.is_tensor(ori_img) else torch.from_numpy(np.ascontiguousarray(ori_img))
            with torch.cuda.device(self.device), torch.cuda.stream(self._p_copy):
                self._p_copy.wait_event(self._p_img_free[ri])         # frame k-3's embedding has read this buffer
                if self._p_img[ri] is None or tuple(self._p_img[ri].shape) != tuple(src.shape):
                    self._p_img[ri] = torch.empty(tuple(src.shape), dtype=torch.uint8, device=self.device)
                self._p_img[ri].copy_(src, non_blocking=True)
                self._p_img_ready[ri].record(self._p_copy)
        with torch.cuda.device(self.device), torch.cuda.stream(pst):
            pst.wait_event(self._p_assoc_done[slot])                  # slot free (frame k-2 associated)
            if n:
                if d.is_cuda:
                    self._after_producer(d, pst, caller)
                    self._p_dets[slot][:n].copy_(d, non_blocking=True)
                else:
                    if self._p_lag > 1 and k >= 2:
                        self._p_embed_done[slot].synchronize()        # frame k-2's copy out of the pinned rows is over
                    self._p_dets_pin[slot][:n].copy_(d)
                    self._p_dets[slot][:n].copy_(self._p_dets_pin[slot][:n], non_blocking=True)
            img_dev = None
            if n:
                if torch.is_tensor(ori_img) and ori_img.is_cuda:
                    self._after_producer(ori_img, pst, caller)
                    img_dev = ori_img.contiguous()
                else:
                    pst.wait_event(self._p_img_ready[ri])
                    img_dev = self._p_img[ri]
            trace = getattr(self, "pipeline_trace", None)       # tools/pipe_trace.py: a list collects timing events
            if trace is not None:
                tev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                tev[0].record(pst)
            _lib.check(self._lib.ssb_embed(self._h, slot, _lib.ptr(self._p_dets[slot]), n,
                                           _lib.ptr(img_dev) if img_dev is not None else None, H, W, 3 * W,
                                           C.c_void_p(pst.cuda_stream)), "ssb_embed")
            self._p_embed_done[slot].record(pst)
            if trace is not None:
                tev[1].record(pst)
            if host_img:
                self._p_img_free[ri].record(pst)
        # The association of THIS frame is enqueued before any earlier frame's result is read back: the host never
        # holds the next ReID launch hostage to a device->host round trip.  track_hint only sizes grids and shared
        # memory, so a bound suffices: the live tracks last read back + the detections of the frames submitted since
        # (every detection can start at most one track).
        hint = min(self.cfg.max_tracks, self._track_hint + sum(self._p_n.values()))
        r = k % self._p_ring
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self.stream.wait_event(self._p_embed_done[slot])
            if trace is not None:
                tev[2].record(self.stream)
            _lib.check(self._lib.ssb_associate(
                self._h, slot, n, H, W, None, C.c_void_p(self._p_out[r].data_ptr() + _HDR_BYTES),
                _lib.ptr(self._p_out[r]), hint, C.c_void_p(self.stream.cuda_stream)),
                "ssb_associate")
            if trace is not None:
                tev[3].record(self.stream)
                trace.append((k, tev))
            self._p_pin[r].copy_(self._p_out[r], non_blocking=True)
            self._p_assoc_done[slot].record(self.stream)
            self._p_res_done[r].record(self.stream)
        self._p_n[k] = n
        self._p_k = k + 1
        # an earlier frame finishes while this frame's embeddings are computed
        return self._pipe_collect(k - self._p_lag) if k >= self._p_lag else None

    def drain_pipelined(self):
        """Rows of every frame still in flight, oldest first (empty list if none); the pipeline is empty afterwards,
        so the next ``update_pipelined`` calls return None again until ``lag`` frames are in flight."""
        if not hasattr(self, "_pstream"):
            return []
        out = [self._pipe_collect(k) for k in range(self._p_done, self._p_k)]
        self._p_k = self._p_done = 0
        self._p_n = {}
        return out

    def flush_pipelined(self):
        """Rows of the last submitted frame (None if nothing is in flight); drains the pipeline (with ``lag`` > 1 use
        ``drain_pipelined`` to get every outstanding frame's rows)."""
        out = self.drain_pipelined()
        return out[-1] if out else None

    def export_tracks(self):
        """Live track table in list order (== Tracker.tracks of the oracle)."""
        torch = self._torch
        T = self._track_hint
        S, D = self.cfg.max_tracks, self.cfg.feat_dim
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            ints = [torch.zeros(S, dtype=torch.int32, device=self.device) for _ in range(6)]
            mean = torch.zeros((S, 8), dtype=torch.float64, device=self.device)
            cov = torch.zeros((S, 8, 8), dtype=torch.float64, device=self.device)
            feat = torch.zeros((S, D), dtype=torch.float32, device=self.device)
            _lib.check(self._lib.ssb_export_tracks(
                self._h, *[_lib.ptr(t) for t in ints], _lib.ptr(mean), _lib.ptr(cov),
                _lib.ptr(feat), C.c_void_p(self.stream.cuda_stream)), "ssb_export_tracks")
        self.stream.synchronize()
        names = ["track_id", "state", "hits", "age", "tsu", "gallery_len"]
        out = {k: v[:T].cpu().numpy().astype(np.int64) for k, v in zip(names, ints)}
        out["mean"] = mean[:T].cpu().numpy()
        out["cov"] = cov[:T].cpu().numpy()
        out["feat"] = feat[:T].cpu().numpy()
        return out

    def debug_costs(self):
        """(cost_a [rows_a, cols_a], cost_b [rows_b, cols_b]) of the last update."""
        self._need_debug("debug_costs")
        pa, pb, pd = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _lib.check(self._lib.ssb_debug_cost_ptrs(self._h, C.byref(pa), C.byref(pb), C.byref(pd)),
                   "ssb_debug_cost_ptrs")
        self.stream.synchronize()
        dims = _wrap_device(self._torch, pd.value, (4,), "<i4", self.device).cpu().numpy()
        ra, ca, rb, cb = [int(x) for x in dims]
        a = np.zeros((ra, ca), dtype=np.float64)
        b = np.zeros((rb, cb), dtype=np.float64)
        if a.size:
            a = _wrap_device(self._torch, pa.value, (ra, ca), "<f8", self.device).cpu().numpy()
        if b.size:
            b = _wrap_device(self._torch, pb.value, (rb, cb), "<f8", self.device).cpu().numpy()
        return a, b


class _DevView:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}


def _wrap_device(torch, ptr, shape, typestr, device):
    """View raw device memory (owned by the tracker workspace) as a tensor."""
    return torch.as_tensor(_DevView(ptr, shape, typestr), device=device)
