"""YOLO post-process on the GPU: confidence filter + class-aware NMS.

Mirrors what ultralytics runs between the detector head and the tracker for
the reference script: conf 0.3, iou 0.4, agnostic False, max_det 1000
(/root/reference/yolo_multi_model.py:18-21).  Input is the
decoded head tensor ``[1 or none, 4+nc(+extra), A]`` (xywh + class scores +
e.g. 51 pose channels); output rows ``[x1,y1,x2,y2,conf,cls,extra...]``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

DEFAULT_CONF, DEFAULT_IOU, DEFAULT_MAX_DET = 0.3, 0.4, 1000


class YoloNMS:
    def __init__(self, num_classes=80, num_extra=0, max_anchors=25200, device="cuda:0",
                 conf=DEFAULT_CONF, iou=DEFAULT_IOU, max_det=DEFAULT_MAX_DET, agnostic=False):
        torch = _lib.require_cuda()
        self._torch, self._lib = torch, _lib.load()
        self.device = torch.device(device)
        self.nc, self.n_extra, self.max_anchors = int(num_classes), int(num_extra), int(max_anchors)
        self.conf, self.iou, self.max_det, self.agnostic = float(conf), float(iou), int(max_det), bool(agnostic)
        nbytes = int(self._lib.ssb_nms_scratch_bytes(self.max_anchors))
        self._scratch = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
        self._out = torch.zeros((self.max_det, 6 + self.n_extra), dtype=torch.float32, device=self.device)
        self._count = torch.zeros(4, dtype=torch.int32, device=self.device)

    def __call__(self, pred, stream=None):
        """pred: float32 CUDA tensor [C, A] or [1, C, A].  Returns (rows_dev
        [max_det, 6+extra], count_dev [1]) -- both stay on the device."""
        torch = self._torch
        if pred.dim() == 3:
            pred = pred[0]
        C_, A = int(pred.shape[0]), int(pred.shape[1])
        if C_ != 4 + self.nc + self.n_extra or A > self.max_anchors:
            raise ValueError(f"pred shape {tuple(pred.shape)} does not match nc={self.nc}, "
                             f"extra={self.n_extra}, max_anchors={self.max_anchors}")
        pred = pred.contiguous()
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        base = (self._scratch.data_ptr() + 255) & ~255
        _lib.check(self._lib.ssb_yolo_nms(
            _lib.ptr(pred), self.nc, self.n_extra, A, self.conf, self.iou, self.max_det,
            int(self.agnostic), _lib.ptr(self._out), _lib.ptr(self._count), C.c_void_p(base),
            C.c_void_p(st.cuda_stream)), "ssb_yolo_nms")
        return self._out, self._count

    def scale_boxes(self, net_hw, frame_hw, stream=None):
        """ultralytics ``scale_boxes``: the NMS rows (network-input pixels) back to the original frame,
        in place on the device (letterbox gain / padding of ``LetterBox(center=True)``)."""
        g, px, py = letterbox_params(net_hw, frame_hw)
        st = stream if stream is not None else self._torch.cuda.current_stream(self.device)
        _lib.check(self._lib.ssb_yolo_scale_boxes(_lib.ptr(self._out), 6 + self.n_extra, _lib.ptr(self._count),
                                                  self.max_det, g, px, py, int(frame_hw[1]), int(frame_hw[0]),
                                                  C.c_void_p(st.cuda_stream)), "ssb_yolo_scale_boxes")
        return self._out, self._count

    def detect(self, pred):
        """Convenience: synchronous, returns float32 ndarray [M, 6+extra]."""
        out, cnt = self(pred)
        m = int(cnt[0].item())
        return out[:m].cpu().numpy()


def letterbox_params(net_hw, frame_hw):
    """(gain, pad_x, pad_y) of ultralytics' scale_boxes for a frame letterboxed into the network input."""
    gain = min(net_hw[0] / frame_hw[0], net_hw[1] / frame_hw[1])
    pad_x = round((net_hw[1] - frame_hw[1] * gain) / 2 - 0.1)
    pad_y = round((net_hw[0] - frame_hw[0] * gain) / 2 - 0.1)
    return float(gain), float(pad_x), float(pad_y)


class YoloV8Decode:
    """Raw YOLOv8 detect / pose head [4*16 + nc + 3*kpts, A] -> decoded [4 + nc + 3*kpts, A]
    (csrc/decode.cu), the tensor ``YoloNMS`` consumes."""

    def __init__(self, num_classes=80, num_kpts=0, in_h=640, in_w=640, device="cuda:0"):
        torch = _lib.require_cuda()
        self._torch, self._lib = torch, _lib.load()
        self.device = torch.device(device)
        self.nc, self.nk, self.in_h, self.in_w = int(num_classes), int(num_kpts), int(in_h), int(in_w)
        self.A = int(self._lib.ssb_yolo_num_anchors(self.in_h, self.in_w))
        if self.A <= 0:
            raise ValueError("network input size must be a positive multiple of 32")
        self._out = torch.empty((4 + self.nc + 3 * self.nk, self.A), dtype=torch.float32, device=self.device)

    def __call__(self, raw, stream=None):
        torch = self._torch
        if raw.dim() == 3:
            raw = raw[0]
        if tuple(raw.shape) != (64 + self.nc + 3 * self.nk, self.A):
            raise ValueError(f"raw head shape {tuple(raw.shape)} != {(64 + self.nc + 3 * self.nk, self.A)}")
        raw = raw.contiguous()
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        _lib.check(self._lib.ssb_yolo_decode_v8(_lib.ptr(raw), self.nc, self.nk, self.in_h, self.in_w,
                                                _lib.ptr(self._out), C.c_void_p(st.cuda_stream)),
                   "ssb_yolo_decode_v8")
        return self._out


class YoloV8Post:
    """The whole detector post-process of a raw YOLOv8 head in ONE C call / 5 launches (``ssb_yolo_postprocess_v8``):
    DFL decode with the best class per anchor fused, confidence filter + class-aware NMS, scale_boxes back to the
    frame fused into the gather.  ``submit`` enqueues the work of one frame on a stream and returns a ticket;
    ``result`` waits for THAT ticket only and returns the detections [M, 6 (+extras)] (a private device copy), so a
    stream loop can keep the post-process of frame k+1 in flight while frame k is being tracked."""

    def __init__(self, num_classes=80, num_kpts=0, in_h=384, in_w=640, frame_hw=(1080, 1920), device="cuda:0",
                 conf=DEFAULT_CONF, iou=DEFAULT_IOU, max_det=DEFAULT_MAX_DET, agnostic=False, depth=3):
        torch = _lib.require_cuda()
        self._torch, self._lib = torch, _lib.load()
        self.device = torch.device(device)
        self.nc, self.nk, self.in_h, self.in_w = int(num_classes), int(num_kpts), int(in_h), int(in_w)
        self.frame_hw = (int(frame_hw[0]), int(frame_hw[1]))
        self.conf, self.iou, self.max_det, self.agnostic = float(conf), float(iou), int(max_det), bool(agnostic)
        self.A = int(self._lib.ssb_yolo_num_anchors(self.in_h, self.in_w))
        if self.A <= 0:
            raise ValueError("network input size must be a positive multiple of 32")
        self.cols = 6 + 3 * self.nk
        self.gain, self.pad_x, self.pad_y = letterbox_params((self.in_h, self.in_w), self.frame_hw)
        nbytes = int(self._lib.ssb_nms_scratch_bytes(self.A))
        dev = self.device
        self._slots = []
        for _ in range(int(depth)):
            self._slots.append({
                "scratch": torch.empty(nbytes + 256, dtype=torch.uint8, device=dev),
                "pred": torch.empty((4 + self.nc + 3 * self.nk, self.A), dtype=torch.float32, device=dev),
                "out": torch.zeros((self.max_det, self.cols), dtype=torch.float32, device=dev),
                "count": torch.zeros(4, dtype=torch.int32, device=dev),
                "count_pin": torch.zeros(4, dtype=torch.int32).pin_memory(),
                "done": torch.cuda.Event(),
            })
        self._k = 0

    def submit(self, raw, stream=None):
        torch = self._torch
        if raw.dim() == 3:
            raw = raw[0]
        if tuple(raw.shape) != (64 + self.nc + 3 * self.nk, self.A):
            raise ValueError(f"raw head shape {tuple(raw.shape)} != {(64 + self.nc + 3 * self.nk, self.A)}")
        raw = raw.contiguous()
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        sl = self._slots[self._k % len(self._slots)]
        self._k += 1
        base = (sl["scratch"].data_ptr() + 255) & ~255
        with torch.cuda.stream(st):
            _lib.check(self._lib.ssb_yolo_postprocess_v8(
                _lib.ptr(raw), self.nc, self.nk, self.in_h, self.in_w, self.conf, self.iou, self.max_det,
                int(self.agnostic), self.gain, self.pad_x, self.pad_y, self.frame_hw[1], self.frame_hw[0],
                _lib.ptr(sl["pred"]), _lib.ptr(sl["out"]), _lib.ptr(sl["count"]), C.c_void_p(base),
                C.c_void_p(st.cuda_stream)), "ssb_yolo_postprocess_v8")
            sl["count_pin"].copy_(sl["count"], non_blocking=True)
            sl["done"].record(st)
        return sl

    def result(self, ticket):
        """Detections of a submitted frame: waits for its event (not for the stream)."""
        ticket["done"].synchronize()
        m = int(ticket["count_pin"][0])
        return ticket["out"][:m]

    def __call__(self, raw, stream=None):
        return self.result(self.submit(raw, stream))


# yolov5n/s/m/l/x anchors (models/yolov5*.yaml), pixels, P3/8, P4/16, P5/32
V5_ANCHORS = np.asarray([[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]],
                        dtype=np.float32).reshape(3, 3, 2)


class YoloV5Decode:
    """Raw YOLOv5 / v7 head [A, 5 + nc] (A = 3 anchors x 3 levels, 25200 at 640x640) -> [4 + nc, A] with
    score = obj * cls (csrc/decode.cu), the tensor ``YoloNMS`` consumes."""

    def __init__(self, num_classes=80, in_h=640, in_w=640, conf=DEFAULT_CONF, anchors=None, device="cuda:0"):
        torch = _lib.require_cuda()
        self._torch, self._lib = torch, _lib.load()
        self.device = torch.device(device)
        self.nc, self.in_h, self.in_w, self.conf = int(num_classes), int(in_h), int(in_w), float(conf)
        n = int(self._lib.ssb_yolo_num_anchors(self.in_h, self.in_w))
        if n <= 0:
            raise ValueError("network input size must be a positive multiple of 32")
        self.A = 3 * n
        a = V5_ANCHORS if anchors is None else np.asarray(anchors, dtype=np.float32).reshape(3, 3, 2)
        self._anchors = torch.as_tensor(np.ascontiguousarray(a)).to(self.device)
        self._out = torch.empty((4 + self.nc, self.A), dtype=torch.float32, device=self.device)

    def __call__(self, raw, stream=None):
        torch = self._torch
        if raw.dim() == 3:
            raw = raw[0]
        if tuple(raw.shape) != (self.A, 5 + self.nc):
            raise ValueError(f"raw head shape {tuple(raw.shape)} != {(self.A, 5 + self.nc)}")
        raw = raw.contiguous()
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        _lib.check(self._lib.ssb_yolo_decode_v5(_lib.ptr(raw), self.nc, self.in_h, self.in_w, self.conf,
                                                _lib.ptr(self._anchors), _lib.ptr(self._out),
                                                C.c_void_p(st.cuda_stream)), "ssb_yolo_decode_v5")
        return self._out


def synth_raw_head_v8(dets, num_classes, in_h, in_w, rng=None, kpts=None):
    """Synthetic RAW v8 head [64 + nc (+3*K), A] whose decode + NMS gives (close to) ``dets`` [N,6]
    (boxes in network-input pixels): each detection is written into the anchor of the finest level
    whose cell contains its centre and whose ltrb distances fit the 0..15 DFL range, as one-hot-ish
    DFL logits interpolating the fractional distance; all other anchors get background logits."""
    from math import floor, log
    rng = rng or np.random.default_rng(0)
    dets = np.asarray(dets, dtype=np.float32).reshape(-1, 6)
    nk = 0 if kpts is None else int(np.asarray(kpts).shape[1])
    levels, base = [], 0
    for s in (8, 16, 32):
        levels.append((s, in_h // s, in_w // s, base))
        base += (in_h // s) * (in_w // s)
    A = base
    raw = np.zeros((64 + num_classes + 3 * nk, A), dtype=np.float32)
    raw[:64] = rng.normal(0, 0.1, (64, A))
    raw[64:64 + num_classes] = rng.uniform(-6.0, -3.5, (num_classes, A))       # sigmoid < 0.03
    used = set()
    for i, (x1, y1, x2, y2, conf, cls) in enumerate(dets):
        cx, cy = (x1 + x2) / 2, (y1 + y2) / 2
        placed = False
        for s, gh, gw, b0 in levels:
            gx0, gy0 = int(floor(cx / s)), int(floor(cy / s))
            # the cell holding the centre first, then its neighbours inside the box (a free anchor per detection)
            for dgy, dgx in ((0, 0), (0, 1), (0, -1), (1, 0), (-1, 0), (1, 1), (-1, -1), (1, -1), (-1, 1)):
                gx, gy = gx0 + dgx, gy0 + dgy
                if not (0 <= gx < gw and 0 <= gy < gh):
                    continue
                a = b0 + gy * gw + gx
                if a in used:
                    continue
                ax, ay = (gx + 0.5) * s, (gy + 0.5) * s
                d = np.array([ax - x1, ay - y1, x2 - ax, y2 - ay], dtype=np.float64) / s
                if d.min() < 0 or d.max() > 14.5:
                    continue
                used.add(a)
                for side in range(4):
                    j, f = int(floor(d[side])), d[side] - floor(d[side])
                    logit = np.full(16, -12.0)
                    f = min(max(f, 1e-4), 1 - 1e-4)
                    logit[j], logit[j + 1] = log(1 - f), log(f)              # softmax -> (1-f, f)
                    raw[side * 16:(side + 1) * 16, a] = logit
                p = min(max(float(conf), 1e-4), 1 - 1e-4)
                raw[64 + int(cls), a] = log(p / (1 - p))
                if nk:
                    kp = np.asarray(kpts)[i]
                    raw[64 + num_classes + 0:64 + num_classes + 3 * nk:3, a] = (kp[:, 0] / s - (gx + 0.5 - 0.5)) / 2
                    raw[64 + num_classes + 1:64 + num_classes + 3 * nk:3, a] = (kp[:, 1] / s - (gy + 0.5 - 0.5)) / 2
                    raw[64 + num_classes + 2:64 + num_classes + 3 * nk:3, a] = 4.0
                placed = True
                break
            if placed:
                break
    return raw


def synth_head(dets, num_classes=80, num_anchors=8400, rng=None, extra=None, jitter=3):
    """Synthetic decoded head tensor [4+nc(+extra), A] float32 whose NMS output
    is (close to) ``dets`` [N,6]: every detection spawns ``jitter`` overlapping
    anchors with lower scores (suppressed by NMS); the remaining anchors carry
    background scores in U(0, 0.05)  (SURVEY.md 8d "raw-head tensors")."""
    rng = rng or np.random.default_rng(0)
    dets = np.asarray(dets, dtype=np.float32).reshape(-1, 6)
    n_extra = 0 if extra is None else int(np.asarray(extra).shape[1])
    pred = np.zeros((4 + num_classes + n_extra, num_anchors), dtype=np.float32)
    pred[4:4 + num_classes] = rng.uniform(0, 0.05, (num_classes, num_anchors)).astype(np.float32)
    pred[0] = rng.uniform(0, 640, num_anchors); pred[1] = rng.uniform(0, 640, num_anchors)
    pred[2] = rng.uniform(8, 64, num_anchors); pred[3] = rng.uniform(8, 64, num_anchors)
    slots = rng.permutation(num_anchors)[:len(dets) * jitter].reshape(len(dets), jitter)
    for i, d in enumerate(dets):
        x1, y1, x2, y2, conf, cls = d
        for k, a in enumerate(slots[i]):
            dx, dy = (0.0, 0.0) if k == 0 else rng.normal(0, 1.0, 2)
            pred[0, a], pred[1, a] = (x1 + x2) / 2 + dx, (y1 + y2) / 2 + dy
            pred[2, a], pred[3, a] = x2 - x1, y2 - y1
            pred[4 + int(cls), a] = conf if k == 0 else conf * rng.uniform(0.5, 0.95)
            if extra is not None:
                pred[4 + num_classes:, a] = np.asarray(extra)[i]
    return pred
