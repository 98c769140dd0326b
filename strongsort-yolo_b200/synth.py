"""Seeded synthetic video streams (frames + detections) for parity tests and bench.

The reference ships no test videos, weights or fixtures;
SURVEY.md 8(d) fixes the synthetic world used everywhere in this repo:
persistent identities with a fixed random texture, constant-velocity motion
with jitter reflecting at the borders, detections = ground-truth boxes + noise
with a small drop / spurious rate, shuffled every frame.

The frames are what ``process()`` hands to the tracker at
/root/reference/yolo_multi_model.py:41 (one BGR uint8 HxWx3 ndarray per call);
the detections are what a detector + NMS would hand to ``StrongSORT.update``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

BASE_SEED = 20240923

# BASELINE.json configs (C1..C5): frame size, dets/frame, persistent / flicker
CONFIGS = {
    "C1": dict(width=640, height=640, n_persistent=10, n_flicker=0),
    "C2": dict(width=1920, height=1080, n_persistent=100, n_flicker=0),
    "C4": dict(width=3840, height=2160, n_persistent=256, n_flicker=244),
}


def _bilinear_resize_u8(tex, out_h, out_w):
    """Half-pixel-centre bilinear resize of a small uint8 [h,w,3] texture."""
    h, w = tex.shape[:2]
    ys = (np.arange(out_h, dtype=np.float32) + 0.5) * (h / out_h) - 0.5
    xs = (np.arange(out_w, dtype=np.float32) + 0.5) * (w / out_w) - 0.5
    ys = np.clip(ys, 0, h - 1)
    xs = np.clip(xs, 0, w - 1)
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    y1 = np.minimum(y0 + 1, h - 1)
    x1 = np.minimum(x0 + 1, w - 1)
    ly = (ys - y0)[:, None, None]
    lx = (xs - x0)[None, :, None]
    t = tex.astype(np.float32)
    top = t[y0][:, x0] * (1 - lx) + t[y0][:, x1] * lx
    bot = t[y1][:, x0] * (1 - lx) + t[y1][:, x1] * lx
    return np.clip(np.rint(top * (1 - ly) + bot * ly), 0, 255).astype(np.uint8)


@dataclass
class Frame:
    img: np.ndarray          # uint8 [H,W,3] BGR
    dets: np.ndarray         # float32 [N,6] x1,y1,x2,y2,conf,cls
    gt_ids: np.ndarray       # int64 [N] identity behind each det (-1 spurious)


class SyntheticStream:
    """One video stream.  ``next_frame()`` is deterministic in (seed, config)."""

    def __init__(self, width=1920, height=1080, n_persistent=100, n_flicker=0,
                 stream_id=0, seed=BASE_SEED, drop_rate=0.02, spurious_rate=0.01,
                 det_noise=1.0, render=True):
        self.W, self.H = int(width), int(height)
        self.K = int(n_persistent)
        self.n_flicker = int(n_flicker)
        self.rng = np.random.default_rng(seed + stream_id)
        self.drop_rate, self.spurious_rate = drop_rate, spurious_rate
        self.det_noise = det_noise
        self.render = render
        rng = self.rng
        s = self.H / 1080.0
        self.h = rng.uniform(80, 240, self.K) * s
        self.w = self.h * rng.uniform(0.3, 0.5, self.K)
        self.cx = rng.uniform(0.1 * self.W, 0.9 * self.W, self.K)
        self.cy = rng.uniform(0.15 * self.H, 0.85 * self.H, self.K)
        self.vx = rng.normal(0, 2.0 * s, self.K)
        self.vy = rng.normal(0, 2.0 * s, self.K)
        # 16x8 random colour blocks per identity -> smooth distinct patterns
        self.tex = rng.integers(0, 256, (self.K + max(self.n_flicker, 1), 16, 8, 3),
                                dtype=np.uint8)
        self.background = rng.integers(96, 160, (self.H, self.W, 3), dtype=np.uint8)
        # short-lived flickers: (cx, cy, w, h, frames_left, tex)
        self.flick = []
        self._patch_cache = {}
        self.t = 0

    # -- world step ---------------------------------------------------------
    def _step(self):
        rng = self.rng
        s = self.H / 1080.0
        self.cx += self.vx + rng.normal(0, 0.5 * s, self.K)
        self.cy += self.vy + rng.normal(0, 0.5 * s, self.K)
        for c, v, lo, hi in ((self.cx, self.vx, 0.05 * self.W, 0.95 * self.W),
                             (self.cy, self.vy, 0.10 * self.H, 0.90 * self.H)):
            under, over = c < lo, c > hi
            c[under] = 2 * lo - c[under]
            c[over] = 2 * hi - c[over]
            v[under | over] *= -1
        # flickers live <= 2 frames; keep the population at n_flicker
        self.flick = [(a, b, c, d, n - 1, e) for (a, b, c, d, n, e) in self.flick if n > 1]
        while len(self.flick) < self.n_flicker:
            hh = rng.uniform(80, 240) * s
            ww = hh * rng.uniform(0.3, 0.5)
            self.flick.append((rng.uniform(0.1 * self.W, 0.9 * self.W),
                               rng.uniform(0.15 * self.H, 0.85 * self.H),
                               ww, hh, int(rng.integers(1, 3)),
                               self.K + int(rng.integers(0, max(self.n_flicker, 1)))))

    def _paint(self, img, cx, cy, w, h, tex_id):
        x1, y1 = int(round(cx - w / 2)), int(round(cy - h / 2))
        ww, hh = max(int(round(w)), 2), max(int(round(h)), 2)
        key = (tex_id, hh, ww)
        patch = self._patch_cache.get(key)
        if patch is None:
            patch = _bilinear_resize_u8(self.tex[tex_id], hh, ww)
            if tex_id < self.K:      # persistent identities keep their size
                self._patch_cache[key] = patch
        sx1, sy1 = max(x1, 0), max(y1, 0)
        sx2, sy2 = min(x1 + ww, self.W), min(y1 + hh, self.H)
        if sx2 > sx1 and sy2 > sy1:
            img[sy1:sy2, sx1:sx2] = patch[sy1 - y1:sy2 - y1, sx1 - x1:sx2 - x1]

    def next_frame(self) -> Frame:
        self._step()
        rng = self.rng
        boxes, ids = [], []
        img = self.background.copy() if self.render else self.background
        for k in range(self.K):
            if self.render:
                self._paint(img, self.cx[k], self.cy[k], self.w[k], self.h[k], k)
            boxes.append((self.cx[k] - self.w[k] / 2, self.cy[k] - self.h[k] / 2,
                          self.cx[k] + self.w[k] / 2, self.cy[k] + self.h[k] / 2))
            ids.append(k)
        for (fx, fy, fw, fh, _n, te) in self.flick:
            if self.render:
                self._paint(img, fx, fy, fw, fh, te)
            boxes.append((fx - fw / 2, fy - fh / 2, fx + fw / 2, fy + fh / 2))
            ids.append(-2)
        boxes = np.asarray(boxes, dtype=np.float64).reshape(-1, 4)
        ids = np.asarray(ids, dtype=np.int64)
        keep = rng.uniform(size=len(boxes)) >= self.drop_rate
        boxes, ids = boxes[keep], ids[keep]
        boxes = boxes + rng.normal(0, self.det_noise, boxes.shape)
        n_sp = int(rng.binomial(max(len(boxes), 1), self.spurious_rate))
        if n_sp:
            s = self.H / 1080.0
            sh = rng.uniform(80, 240, n_sp) * s
            sw = sh * rng.uniform(0.3, 0.5, n_sp)
            scx = rng.uniform(0.1 * self.W, 0.9 * self.W, n_sp)
            scy = rng.uniform(0.15 * self.H, 0.85 * self.H, n_sp)
            sp = np.stack([scx - sw / 2, scy - sh / 2, scx + sw / 2, scy + sh / 2], 1)
            boxes = np.concatenate([boxes, sp], 0)
            ids = np.concatenate([ids, -np.ones(n_sp, dtype=np.int64)])
        # keep boxes well inside the frame so crops are never degenerate
        boxes[:, 0] = np.clip(boxes[:, 0], 0, self.W - 8)
        boxes[:, 1] = np.clip(boxes[:, 1], 0, self.H - 8)
        boxes[:, 2] = np.clip(boxes[:, 2], boxes[:, 0] + 4, self.W - 1)
        boxes[:, 3] = np.clip(boxes[:, 3], boxes[:, 1] + 4, self.H - 1)
        conf = rng.uniform(0.5, 0.95, len(boxes))
        perm = rng.permutation(len(boxes))
        dets = np.concatenate([boxes, conf[:, None], np.zeros((len(boxes), 1))], 1)
        self.t += 1
        return Frame(img=img, dets=dets[perm].astype(np.float32), gt_ids=ids[perm])


def make_stream(config="C2", stream_id=0, **kw) -> SyntheticStream:
    cfg = dict(CONFIGS[config])
    cfg.update(kw)
    return SyntheticStream(stream_id=stream_id, **cfg)


def calibration_crops(n=192, seed=BASE_SEED - 1):
    """Crops (uint8 frame + int boxes) used once to calibrate the synthetic
    OSNet's BatchNorm statistics (tools/make_osnet_weights.py)."""
    st = SyntheticStream(width=1280, height=720, n_persistent=n, seed=seed,
                         drop_rate=0.0, spurious_rate=0.0)
    fr = st.next_frame()
    return fr.img, fr.dets
