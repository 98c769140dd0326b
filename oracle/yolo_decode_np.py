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


V5_ANCHORS = np.asarray([[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]],
                        dtype=np.float32).reshape(3, 3, 2)


def decode_v5(raw, nc, in_h, in_w, conf_thres, anchors=V5_ANCHORS):
    """YOLOv5 / v7 Detect inference + the scoring of its non_max_suppression (yolov5 models/yolo.py,
    utils/general.py -- third-party, not vendored): raw float32 [A, 5+nc] logits, A = 3
    anchors x the three grids, level by level with [na, ny, nx] order inside a level ->
    float32 [4+nc, A]: xywh pixels and obj*cls scores, zero where obj <= conf_thres."""
    raw = torch.as_tensor(np.asarray(raw, dtype=np.float32))
    out, o = [], 0
    for lvl, s in enumerate((8, 16, 32)):
        ny, nx = in_h // s, in_w // s
        n = 3 * ny * nx
        y = raw[o:o + n].view(3, ny, nx, 5 + nc).sigmoid()
        o += n
        yv, xv = torch.meshgrid(torch.arange(ny, dtype=torch.float32), torch.arange(nx, dtype=torch.float32), indexing="ij")
        grid = torch.stack((xv, yv), -1).view(1, ny, nx, 2)
        xy = (y[..., 0:2] * 2.0 - 0.5 + grid) * float(s)
        wh = (y[..., 2:4] * 2.0) ** 2 * torch.as_tensor(anchors[lvl]).view(3, 1, 1, 2)
        obj = y[..., 4:5]
        sc = torch.where(obj > conf_thres, y[..., 5:] * obj, torch.zeros(()))
        out.append(torch.cat((xy, wh, sc), -1).view(n, 4 + nc))
    return torch.cat(out, 0).T.contiguous().numpy().astype(np.float32)
