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

    def detect(self, pred):
        """Convenience: synchronous, returns float32 ndarray [M, 6+extra]."""
        out, cnt = self(pred)
        m = int(cnt[0].item())
        return out[:m].cpu().numpy()


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
