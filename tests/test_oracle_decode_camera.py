"""CPU: the YOLOv8 head-decode restatement inverts the synthetic raw head, and the oracle's
camera_update behaves like upstream's corner warp."""
import numpy as np

from oracle import nms_np, strongsort_np as ss, yolo_decode_np
from strongsort_yolo_b200 import yolo


def _dets(rng, n, in_h, in_w):
    cx = rng.uniform(60, in_w - 60, n); cy = rng.uniform(60, in_h - 60, n)
    w = rng.uniform(20, 90, n); h = rng.uniform(30, 110, n)
    d = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2, rng.uniform(0.5, 0.95, n),
                  rng.integers(0, 3, n).astype(np.float64)], 1)
    return d.astype(np.float32)


def test_decode_inverts_synthetic_raw_head():
    rng = np.random.default_rng(5)
    in_h, in_w, nc = 384, 640, 80                      # 1080p letterboxed: A = 5040
    dets = _dets(rng, 12, in_h, in_w)
    raw = yolo.synth_raw_head_v8(dets, nc, in_h, in_w, rng=rng)
    assert raw.shape == (64 + nc, 5040)
    pred = yolo_decode_np.decode_v8(raw, nc, 0, in_h, in_w)
    rows = nms_np.yolo_nms(pred, nc, 0, 0.3, 0.4, 1000, False)
    assert len(rows) == len(dets)
    order = np.argsort(-dets[:, 4], kind="stable")
    np.testing.assert_allclose(rows[:, :4], dets[order, :4], atol=2e-2)
    np.testing.assert_allclose(rows[:, 4], dets[order, 4], atol=1e-4)
    np.testing.assert_array_equal(rows[:, 5], dets[order, 5])


def test_decode_pose_channels():
    rng = np.random.default_rng(6)
    in_h = in_w = 640
    dets = _dets(rng, 5, in_h, in_w)
    dets[:, 5] = 0                                      # single-class pose head
    dets[:, [0, 2]] = dets[:, [0, 2]] - dets[:, [0]] + 20 + 120 * np.arange(5)[:, None]   # no overlaps
    kp = np.stack([rng.uniform(100, 500, (5, 17)), rng.uniform(100, 500, (5, 17)), np.ones((5, 17))], 2)
    raw = yolo.synth_raw_head_v8(dets, 1, in_h, in_w, rng=rng, kpts=kp)
    pred = yolo_decode_np.decode_v8(raw, 1, 17, in_h, in_w)
    rows = nms_np.yolo_nms(pred, 1, 51, 0.3, 0.4, 1000, False)
    assert rows.shape == (5, 6 + 51)
    order = np.argsort(-dets[:, 4], kind="stable")
    got = rows[:, 6:].reshape(5, 17, 3)
    np.testing.assert_allclose(got[:, :, :2], kp[order][:, :, :2], atol=1e-2)
    assert (got[:, :, 2] > 0.9).all()


def test_oracle_camera_update_identity_and_shift():
    ora = ss.StrongSORTOracle(None)
    rng = np.random.default_rng(0)
    dets = _dets(rng, 6, 480, 640)
    feats = np.maximum(rng.normal(0, 1, (6, 512)), 0).astype(np.float32)
    img = np.zeros((480, 640, 3), dtype=np.uint8)
    for _ in range(3):
        ora.update(dets, img, features=feats)
    before = np.stack([t.mean.copy() for t in ora.tracker.tracks])
    ora.tracker.camera_update(np.eye(2, 3))
    np.testing.assert_allclose(np.stack([t.mean for t in ora.tracker.tracks]), before, rtol=1e-14)
    ora.tracker.camera_update(np.array([[1, 0, 5.0], [0, 1, -3.0]]))
    after = np.stack([t.mean for t in ora.tracker.tracks])
    np.testing.assert_allclose(after[:, 0], before[:, 0] + 5.0, rtol=1e-12)
    np.testing.assert_allclose(after[:, 1], before[:, 1] - 3.0, rtol=1e-12)
    np.testing.assert_allclose(after[:, 2:], before[:, 2:], rtol=1e-12)


def test_ecc_warp_estimates_translation():
    """Host-side helper of StrongSORT.camera_update (cv2.findTransformECC, upstream's settings)."""
    import pytest
    cv2 = pytest.importorskip("cv2")
    from strongsort_yolo_b200 import ecc
    rng = np.random.default_rng(0)
    coarse = rng.uniform(0, 255, (12, 16, 3)).astype(np.float32)
    base = np.clip(cv2.resize(coarse, (640, 480), interpolation=cv2.INTER_CUBIC), 0, 255).astype(np.uint8)
    m = np.float32([[1, 0, 20], [0, 1, -10]])
    moved = cv2.warpAffine(base, m, (640, 480), borderMode=cv2.BORDER_REFLECT)
    w = ecc.ecc_warp(base, moved)
    assert w is not None and w.shape == (2, 3)
    assert abs(abs(w[0, 2]) - 20) < 3 and abs(abs(w[1, 2]) - 10) < 3
    assert abs(w[0, 0] - 1) < 0.02 and abs(w[0, 1]) < 0.02


def test_decode_v5_restatement_shapes_and_rule():
    rng = np.random.default_rng(12)
    in_h = in_w = 64                                   # A = 3 * (64 + 16 + 4) = 252
    raw = rng.normal(0, 2.0, (252, 5 + 4)).astype(np.float32)
    out = yolo_decode_np.decode_v5(raw, 4, in_h, in_w, 0.3)
    assert out.shape == (8, 252)
    obj = 1 / (1 + np.exp(-raw[:, 4]))
    assert (out[4:, obj <= 0.3] == 0).all() and (out[4:, obj > 0.3] > 0).all()
    # first anchor of the first level, cell (0, 0): xy = (2 s - 0.5) * 8, wh = (2 s)^2 * (10, 13)
    s = 1 / (1 + np.exp(-raw[0, :4].astype(np.float64)))
    np.testing.assert_allclose(out[:4, 0], [(2 * s[0] - 0.5) * 8, (2 * s[1] - 0.5) * 8, (2 * s[2]) ** 2 * 10,
                                            (2 * s[3]) ** 2 * 13], rtol=1e-5)
