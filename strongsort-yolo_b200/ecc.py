"""Camera-motion estimate for ``StrongSORT.camera_update``: upstream's ``ECC()``
helper around ``cv2.findTransformECC`` -- grayscale, 0.1x downscale, MOTION_EUCLIDEAN, eps 1e-5,
100 iterations, translation scaled back to full resolution.  Host-side and optional (outside
``update()``); cv2 is imported lazily and its absence raises.
"""
from __future__ import annotations

import numpy as np


def ecc_warp(src, dst, scale=0.1, eps=1e-5, max_iter=100):
    """-> float64 [2,3] warp aligning ``src`` (previous frame) to ``dst``, or None if ECC fails."""
    import cv2
    assert src.shape == dst.shape, "the source image must be the same format to the target image!"
    if src.ndim == 3:
        src = cv2.cvtColor(src, cv2.COLOR_BGR2GRAY)
        dst = cv2.cvtColor(dst, cv2.COLOR_BGR2GRAY)
    if scale is not None and scale != 1:
        src_r = cv2.resize(src, (0, 0), fx=scale, fy=scale, interpolation=cv2.INTER_LINEAR)
        dst_r = cv2.resize(dst, (0, 0), fx=scale, fy=scale, interpolation=cv2.INTER_LINEAR)
        sc = [scale, scale]
    else:
        src_r, dst_r, sc = src, dst, None
    warp = np.eye(2, 3, dtype=np.float32)
    criteria = (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, max_iter, eps)
    try:
        _, warp = cv2.findTransformECC(src_r, dst_r, warp, cv2.MOTION_EUCLIDEAN, criteria, None, 1)
    except cv2.error:
        return None
    if sc is not None:
        warp[0, 2] = warp[0, 2] / sc[0]
        warp[1, 2] = warp[1, 2] / sc[1]
    return warp.astype(np.float64)
