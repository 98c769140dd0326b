"""YOLOv8 Detect / Pose head decode, CPU restatement (TEST INFRASTRUCTURE; parity UNPINNED --
ultralytics is not installed and not vendored).  Restates the inference branch of
ultralytics.nn.modules.head.Detect / Pose: DFL (softmax over reg_max=16 bins, expectation),
``dist2bbox(..., xywh=True)`` around ``make_anchors(..., offset=0.5)`` anchor points, times the
stride; ``cls.sigmoid()``; keypoints ``(v * 2 + (anchor - 0.5)) * stride``, visibility sigmoid.
"""
import numpy as np
import torch

REG_MAX = 16


def make_anchors(in_h, in_w, strides=(8, 16, 32)):
    pts, st = [], []
    for s in strides:
        h, w = in_h // s, in_w // s
        sx = torch.arange(w, dtype=torch.float32) + 0.5
        sy = torch.arange(h, dtype=torch.float32) + 0.5
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s), dtype=torch.float32))
    return torch.cat(pts).T, torch.cat(st).T          # [2, A], [1, A]


def decode_v8(raw, nc, nk, in_h, in_w):
    """raw float32 [64 + nc + 3*nk, A] -> float32 [4 + nc + 3*nk, A]."""
    raw = torch.as_tensor(np.asarray(raw, dtype=np.float32))
    A = raw.shape[1]
    anchors, strides = make_anchors(in_h, in_w)
    assert anchors.shape[1] == A
    box, cls, kpt = raw[:64], raw[64:64 + nc], raw[64 + nc:]
    proj = torch.arange(REG_MAX, dtype=torch.float32)
    dist = (box.view(4, REG_MAX, A).softmax(1) * proj.view(1, REG_MAX, 1)).sum(1)      # [4, A]
    lt, rb = dist[:2], dist[2:]
    x1y1, x2y2 = anchors - lt, anchors + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 0) * strides
    out = [dbox, cls.sigmoid()]
    if nk:
        k = kpt.view(nk, 3, A).clone()
        k[:, 0] = (k[:, 0] * 2.0 + (anchors[0] - 0.5)) * strides[0]
        k[:, 1] = (k[:, 1] * 2.0 + (anchors[1] - 0.5)) * strides[0]
        k[:, 2] = k[:, 2].sigmoid()
        out.append(k.view(nk * 3, A))
    return torch.cat(out, 0).numpy().astype(np.float32)
