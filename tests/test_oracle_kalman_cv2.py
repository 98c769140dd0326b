"""CPU: the oracle's Kalman filter (constant-velocity model on (cx, cy, a, h), NSA
measurement noise, Cholesky-solve update) against an INSTALLED independent implementation of the
same equations -- OpenCV's cv2.KalmanFilter in float64 -- fed with the oracle's own F, H, Q, R.
Pins the predict / update algebra the way tests/test_oracle_lsap.py pins the assignment step against
scipy; the noise schedules themselves (which Q and R) remain the recalled upstream constants."""
import numpy as np
import pytest

from oracle import strongsort_np as ss

cv2 = pytest.importorskip("cv2")


def _cv_filter(mean, cov, Q, R):
    kf = cv2.KalmanFilter(8, 4, 0, cv2.CV_64F)
    F = np.eye(8)
    for i in range(4):
        F[i, 4 + i] = 1.0
    H = np.eye(4, 8)
    kf.transitionMatrix = F
    kf.measurementMatrix = H
    kf.processNoiseCov = Q
    kf.measurementNoiseCov = R
    kf.statePost = mean.reshape(8, 1).copy()
    kf.errorCovPost = cov.copy()
    return kf


@pytest.mark.parametrize("conf", [0.0, 0.35, 0.9])
def test_predict_update_match_cv2(conf):
    rng = np.random.default_rng(int(conf * 100))
    kf = ss.KalmanFilter()
    for _ in range(20):
        z0 = np.array([rng.uniform(50, 1800), rng.uniform(50, 1000), rng.uniform(0.3, 0.5), rng.uniform(80, 240)],
                      dtype=np.float32)
        mean, cov = kf.initiate(z0)
        for step in range(4):
            wp, wv = kf._std_weight_position, kf._std_weight_velocity
            std = np.array([wp * mean[0], wp * mean[1], 1 * mean[2], wp * mean[3],
                            wv * mean[0], wv * mean[1], 0.1 * mean[2], wv * mean[3]])
            Q = np.diag(np.square(std))
            ref = _cv_filter(mean, cov, Q, np.eye(4))
            ref.predict()
            mean, cov = kf.predict(mean, cov)
            np.testing.assert_allclose(mean, ref.statePre.ravel(), rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(cov, ref.errorCovPre, rtol=1e-11, atol=1e-11)
            # update with a noisy measurement; R = the NSA noise of project(mean, cov, conf)
            z = (mean[:4] + rng.normal(0, [2.0, 2.0, 0.01, 2.0])).astype(np.float32)
            r = np.array([wp * mean[3], wp * mean[3], 1e-1, wp * mean[3]]) * (1.0 - conf)
            ref.measurementNoiseCov = np.diag(np.square(r))
            ref.correct(z.astype(np.float64).reshape(4, 1))
            mean, cov = kf.update(mean, cov, z, conf)
            np.testing.assert_allclose(mean, ref.statePost.ravel(), rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(cov, ref.errorCovPost, rtol=1e-8, atol=1e-8)


def test_gating_distance_is_mahalanobis():
    """gating_distance == (z - Hm)^T S^-1 (z - Hm) with S = H P H^T + R(conf = 0), via numpy.linalg.inv."""
    rng = np.random.default_rng(5)
    kf = ss.KalmanFilter()
    mean, cov = kf.initiate(np.array([400.0, 300.0, 0.4, 150.0], dtype=np.float32))
    for _ in range(3):
        mean, cov = kf.predict(mean, cov)
    Z = (mean[:4] + rng.normal(0, [5, 5, 0.02, 5], (12, 4))).astype(np.float32)
    d = kf.gating_distance(mean, cov, Z)
    mu, S = kf.project(mean, cov)
    ref = np.array([(z - mu) @ np.linalg.inv(S) @ (z - mu) for z in Z.astype(np.float64)])
    np.testing.assert_allclose(d, ref, rtol=1e-9)


def test_iou_matches_torchvision_box_iou():
    """iou_matching.iou (A.7) against the installed torchvision.ops.box_iou on the same boxes."""
    import torch
    import torchvision
    rng = np.random.default_rng(1)
    tl = rng.uniform(0, 800, (40, 2)); wh = rng.uniform(20, 200, (40, 2))
    cand = np.concatenate([tl, wh], 1).astype(np.float32)                 # tlwh
    for k in range(5):
        bbox = np.concatenate([rng.uniform(0, 800, 2), rng.uniform(20, 200, 2)])
        got = ss.iou(bbox, cand)
        a = torch.tensor([[bbox[0], bbox[1], bbox[0] + bbox[2], bbox[1] + bbox[3]]], dtype=torch.float64)
        c = torch.as_tensor(cand.astype(np.float64))
        b = torch.cat([c[:, :2], c[:, :2] + c[:, 2:]], 1)
        ref = torchvision.ops.box_iou(a, b)[0].numpy()
        np.testing.assert_allclose(got, ref, rtol=0, atol=2e-6)             # float32 corner/area terms in the oracle


def test_nn_cosine_distance_matches_scipy_cdist():
    """NearestNeighborDistanceMetric's cosine metric (A.5) against scipy.spatial.distance.cdist."""
    from scipy.spatial.distance import cdist
    rng = np.random.default_rng(2)
    gallery = np.maximum(rng.normal(0, 1, (37, 512)), 0).astype(np.float32)
    feats = np.maximum(rng.normal(0, 1, (23, 512)), 0).astype(np.float32)
    got = ss._nn_cosine_distance(gallery, feats)
    ref = cdist(gallery.astype(np.float64), feats.astype(np.float64), metric="cosine").min(axis=0)
    np.testing.assert_allclose(got, ref, rtol=0, atol=5e-6)
