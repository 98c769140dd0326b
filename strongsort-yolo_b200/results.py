"""Result objects shaped the way the reference's ``process()`` consumes them
(/root/reference/yolo_multi_model.py:45-162): ``r.boxes`` is iterable and has
``.id`` (None => caller skips the frame, :54); every box exposes 1-element
``.conf/.cls/.xyxy/.id`` sequences (``zip`` at :73/:126, ``int(bbox.id)`` at :46);
``r.keypoints`` is None or iterable with ``.xy.tolist()`` (:58-62); ``r.masks``
is None or iterable with ``.xy`` polygons (:71-72); ``r.names[int(cls)]`` (:86).
"""
from __future__ import annotations

import numpy as np


class Boxes:
    def __init__(self, xyxy, conf, cls, ids=None):
        self.xyxy = np.asarray(xyxy, dtype=np.float32).reshape(-1, 4)
        self.conf = np.asarray(conf, dtype=np.float32).reshape(-1)
        self.cls = np.asarray(cls, dtype=np.float32).reshape(-1)
        self.id = None if ids is None else np.asarray(ids, dtype=np.int64).reshape(-1)

    def __len__(self):
        return len(self.conf)

    def __getitem__(self, i):
        i = slice(i, i + 1) if isinstance(i, (int, np.integer)) else i
        return Boxes(self.xyxy[i], self.conf[i], self.cls[i], None if self.id is None else self.id[i])

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


class _XY:
    def __init__(self, xy):
        self.xy = xy


class Keypoints:
    """[M,17,2] (+ visibility); iterating yields objects with ``.xy`` of shape [1,17,2]."""

    def __init__(self, kpts):
        k = np.asarray(kpts, dtype=np.float32)
        self.data = k
        self.xy = k[..., :2]

    def __len__(self):
        return len(self.data)

    def __iter__(self):
        for i in range(len(self)):
            yield _XY(self.xy[i:i + 1])


class Masks:
    def __init__(self, polygons):
        self.xy = list(polygons)

    def __len__(self):
        return len(self.xy)

    def __iter__(self):
        for p in self.xy:
            yield _XY([p])


class Results:
    def __init__(self, boxes, names, keypoints=None, masks=None, orig_shape=None):
        self.boxes, self.names = boxes, names
        self.keypoints, self.masks = keypoints, masks
        self.orig_shape = orig_shape


DEFAULT_NAMES = {0: "person"}


def results_from_tracks(rows, det_index, names=None, keypoints=None, masks=None, orig_shape=None):
    """rows: [M,7] from ``StrongSORT.update``; det_index: ``tracker.last_det_index``
    (source detection of every row, -1 when coasting).  Per-detection extras
    (keypoints [N,17,2or3], mask polygons) are re-indexed to follow their boxes,
    as ultralytics does with the trailing idx column; rows
    without a detection this frame get zeros (drawn as invalid, :60-62)."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, 7)
    names = names or DEFAULT_NAMES
    if len(rows) == 0:
        return Results(Boxes(np.zeros((0, 4)), [], [], None), names, None, None, orig_shape)
    det_index = np.asarray(det_index, dtype=np.int64).reshape(-1)
    boxes = Boxes(rows[:, :4], rows[:, 6], rows[:, 5], rows[:, 4])
    kp = None
    if keypoints is not None:
        k = np.asarray(keypoints, dtype=np.float32)
        out = np.zeros((len(rows),) + k.shape[1:], dtype=np.float32)
        ok = det_index >= 0
        out[ok] = k[det_index[ok]]
        kp = Keypoints(out)
    mk = None
    if masks is not None:
        mk = Masks([masks[i] if i >= 0 else np.zeros((0, 2), np.float32) for i in det_index])
    return Results(boxes, names, kp, mk, orig_shape)


def label_lines(results, frame_id=0):
    """The labels-file lines of /root/reference/yolo_multi_model.py:165-169:
    ``frameId cls id conf x1 y1 x2 y2 -1 -1 -1 -1`` with int-truncated coords
    (the reference always writes frameId 0, :32 -- kept as a parameter)."""
    out = []
    b = results.boxes
    if b is None or b.id is None:
        return out
    for xyxy, conf, cls, tid in zip(b.xyxy, b.conf, b.cls, b.id):
        out.append(f"{frame_id} {int(cls)} {int(tid)} {round(float(conf), 3)} {int(xyxy[0])} "
                   f"{int(xyxy[1])} {int(xyxy[2])} {int(xyxy[3])} -1 -1 -1 -1\n")
    return out
