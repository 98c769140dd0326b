"""YOLO confidence filter + class-aware NMS, CPU restatement (TEST INFRASTRUCTURE).

PARITY PINNED AGAINST THE INSTALLED LIBRARY for the suppression step:
torchvision.ops.nms (0.26, installed) is called directly; the surrounding
logic restates ultralytics' ``non_max_suppression`` (third-party, unpinned,
not installed -- SURVEY.md C.2) with the reference's thresholds
(/root/reference/yolo_multi_model.py:18-21): best class only, candidates
``max_cls_conf > conf``, xywh->xyxy in float32, class offset 7680 unless
agnostic, keep the first ``max_det`` survivors in descending-score order.
"""
import numpy as np
import torch
import torchvision

MAX_WH = 7680.0


def yolo_nms(pred, nc, n_extra, conf_thres, iou_thres, max_det, agnostic):
    """pred float32 [4+nc+extra, A] -> float32 [M, 6+extra]."""
    p = torch.as_tensor(np.asarray(pred, dtype=np.float32)).T          # [A, C]
    box, cls, extra = p[:, :4], p[:, 4:4 + nc], p[:, 4 + nc:4 + nc + n_extra]
    conf, j = cls.max(1, keepdim=True)
    keep = conf.view(-1) > conf_thres
    dw, dh = box[:, 2] / 2, box[:, 3] / 2
    xyxy = torch.stack([box[:, 0] - dw, box[:, 1] - dh, box[:, 0] + dw, box[:, 1] + dh], 1)
    x = torch.cat([xyxy, conf, j.float(), extra], 1)[keep]
    if x.shape[0] == 0:
        return np.zeros((0, 6 + n_extra), dtype=np.float32)
    c = x[:, 5:6] * (0.0 if agnostic else MAX_WH)
    i = torchvision.ops.nms(x[:, :4] + c, x[:, 4], iou_thres)[:max_det]
    return x[i].numpy().astype(np.float32)


def scale_boxes(rows, net_hw, frame_hw):
    """ultralytics ops.scale_boxes + clip_boxes restated (third-party, not vendored): NMS rows in
    network-input pixels -> original frame pixels (float32 arithmetic, in the library's order)."""
    rows = np.array(rows, dtype=np.float32, copy=True)
    gain = min(net_hw[0] / frame_hw[0], net_hw[1] / frame_hw[1])
    pad_x = round((net_hw[1] - frame_hw[1] * gain) / 2 - 0.1)
    pad_y = round((net_hw[0] - frame_hw[0] * gain) / 2 - 0.1)
    b = torch.as_tensor(rows[:, :4])
    b[:, 0] -= pad_x; b[:, 2] -= pad_x
    b[:, 1] -= pad_y; b[:, 3] -= pad_y
    b /= gain
    b[:, 0].clamp_(0, frame_hw[1]); b[:, 2].clamp_(0, frame_hw[1])
    b[:, 1].clamp_(0, frame_hw[0]); b[:, 3].clamp_(0, frame_hw[0])
    rows[:, :4] = b.numpy()
    return rows
