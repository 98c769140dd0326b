"""NumPy/SciPy restatement of the StrongSORT association path (TEST INFRASTRUCTURE).

PARITY UNPINNED -- the reference tree has no StrongSORT code (the only tracker
call is ``model.track(... tracker="botsort.yaml")`` at
/root/reference/yolo_multi_model.py:41).  This file restates the algorithm that
upstream vendored under ``yolov5/strong_sort`` / ``yolov7/strong_sort`` (both
directories are empty in the snapshot), following SURVEY.md Appendix A:

  A.1 defaults          -> ``StrongSORTOracle.__init__``
  A.2 update(dets,img)  -> ``StrongSORTOracle.update``
  A.4 Kalman filter     -> ``KalmanFilter``
  A.5 cosine NN metric  -> ``NearestNeighborDistanceMetric``
  A.6 association       -> ``min_cost_matching`` / ``Tracker._match``
  A.7 IoU cost          -> ``iou`` / ``iou_cost``
  A.8 Track lifecycle   -> ``Track``

Every dtype is written out (NumPy 1.x-era value-based casting and NumPy 2's
NEP-50 rules disagree on a few scalar promotions; the choices pinned here are
listed in DESIGN.md "oracle pins").  The CUDA path must reproduce the same
decisions (assignment indices, track IDs) and floats within 1e-3 rel.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
from scipy.optimize import linear_sum_assignment

INFTY_COST = 1e5
# chi-square 0.95 quantiles, N degrees of freedom
chi2inv95 = {1: 3.8415, 2: 5.9915, 3: 7.8147, 4: 9.4877, 5: 11.070,
             6: 12.592, 7: 14.067, 8: 15.507, 9: 16.919}

TENTATIVE, CONFIRMED, DELETED = 1, 2, 3


# ----------------------------------------------------------------------------
# A.4 Kalman filter: state [cx, cy, a, h, vcx, vcy, va, vh], float64
# ----------------------------------------------------------------------------
class KalmanFilter:
    ndim = 4

    def __init__(self):
        dt = 1.0
        self._motion_mat = np.eye(8, 8, dtype=np.float64)
        for i in range(4):
            self._motion_mat[i, 4 + i] = dt
        self._update_mat = np.eye(4, 8, dtype=np.float64)
        self._std_weight_position = 1.0 / 20
        self._std_weight_velocity = 1.0 / 160

    def initiate(self, measurement):
        # pin: measurement arrives as float32 xyah (Detection.to_xyah); it is
        # widened to float64 exactly, std is evaluated in float64.
        z = np.asarray(measurement, dtype=np.float64)
        mean = np.r_[z, np.zeros(4, dtype=np.float64)]
        wp, wv = self._std_weight_position, self._std_weight_velocity
        std = np.array([2 * wp * z[0], 2 * wp * z[1], 1 * z[2], 2 * wp * z[3],
                        10 * wv * z[0], 10 * wv * z[1], 0.1 * z[2], 10 * wv * z[3]],
                       dtype=np.float64)
        covariance = np.diag(np.square(std))
        return mean, covariance

    def predict(self, mean, covariance):
        wp, wv = self._std_weight_position, self._std_weight_velocity
        std = np.array([wp * mean[0], wp * mean[1], 1 * mean[2], wp * mean[3],
                        wv * mean[0], wv * mean[1], 0.1 * mean[2], wv * mean[3]],
                       dtype=np.float64)
        motion_cov = np.diag(np.square(std))
        mean = np.dot(self._motion_mat, mean)
        covariance = np.linalg.multi_dot(
            (self._motion_mat, covariance, self._motion_mat.T)) + motion_cov
        return mean, covariance

    def project(self, mean, covariance, confidence=0.0):
        wp = self._std_weight_position
        std = [wp * mean[3], wp * mean[3], 1e-1, wp * mean[3]]
        std = np.array([(1.0 - confidence) * x for x in std], dtype=np.float64)
        innovation_cov = np.diag(np.square(std))
        mean = np.dot(self._update_mat, mean)
        covariance = np.linalg.multi_dot(
            (self._update_mat, covariance, self._update_mat.T))
        return mean, covariance + innovation_cov

    def update(self, mean, covariance, measurement, confidence=0.0):
        projected_mean, projected_cov = self.project(mean, covariance, confidence)
        chol_factor, lower = scipy.linalg.cho_factor(
            projected_cov, lower=True, check_finite=False)
        kalman_gain = scipy.linalg.cho_solve(
            (chol_factor, lower), np.dot(covariance, self._update_mat.T).T,
            check_finite=False).T
        innovation = np.asarray(measurement, dtype=np.float64) - projected_mean
        new_mean = mean + np.dot(innovation, kalman_gain.T)
        new_covariance = covariance - np.linalg.multi_dot(
            (kalman_gain, projected_cov, kalman_gain.T))
        return new_mean, new_covariance

    def gating_distance(self, mean, covariance, measurements):
        """Squared Mahalanobis distance (4 dof) of every measurement row."""
        mean, covariance = self.project(mean, covariance)  # confidence = 0
        cholesky_factor = np.linalg.cholesky(covariance)
        d = np.asarray(measurements, dtype=np.float64) - mean
        z = scipy.linalg.solve_triangular(
            cholesky_factor, d.T, lower=True, check_finite=False)
        return np.sum(z * z, axis=0)


# ----------------------------------------------------------------------------
# Detection / Track (A.8)
# ----------------------------------------------------------------------------
class Detection:
    def __init__(self, tlwh, confidence, feature):
        self.tlwh = np.asarray(tlwh, dtype=np.float32)
        self.confidence = float(confidence)
        self.feature = np.asarray(feature, dtype=np.float32)

    def to_xyah(self):
        # float32 arithmetic, as the array ops on a float32 tlwh are upstream
        ret = self.tlwh.copy()
        ret[:2] += ret[2:] / np.float32(2)
        ret[2] /= ret[3]
        return ret


def _l2norm32(v):
    """float32 L2 norm restated without BLAS: sqrt(sum(v*v)) accumulated in
    float64 then rounded -- within 1 ulp of np.linalg.norm on float32."""
    v = np.asarray(v, dtype=np.float32)
    return np.float32(np.sqrt(np.sum(v.astype(np.float64) ** 2)))


class Track:
    def __init__(self, mean, covariance, track_id, class_id, conf, n_init,
                 max_age, ema_alpha, feature):
        self.mean = mean
        self.covariance = covariance
        self.track_id = track_id
        self.class_id = class_id
        self.conf = conf
        self.hits = 1
        self.age = 1
        self.time_since_update = 0
        self.ema_alpha = ema_alpha
        self.state = TENTATIVE
        self.features = []
        if feature is not None:
            feature = np.asarray(feature, dtype=np.float32)
            feature = feature / _l2norm32(feature)
            self.features.append(feature)
        self._n_init = n_init
        self._max_age = max_age
        self.kf = KalmanFilter()

    def to_tlwh(self):
        ret = self.mean[:4].copy()
        ret[2] *= ret[3]
        ret[:2] -= ret[2:] / 2
        return ret

    def predict(self):
        self.mean, self.covariance = self.kf.predict(self.mean, self.covariance)
        self.age += 1
        self.time_since_update += 1

    def camera_update(self, warp_matrix):
        """Upstream Track.camera_update after its ECC call: warp the tl / br corners
        with the 2x3 matrix (homogeneous row appended) and rewrite mean[:4].  float64."""
        m = np.vstack([np.asarray(warp_matrix, dtype=np.float64).reshape(2, 3), [0.0, 0.0, 1.0]])
        tlbr = self.to_tlwh()
        tlbr[2:] = tlbr[:2] + tlbr[2:]
        x1, y1, x2, y2 = tlbr
        x1_ = m[0, 0] * x1 + m[0, 1] * y1 + m[0, 2]
        y1_ = m[1, 0] * x1 + m[1, 1] * y1 + m[1, 2]
        x2_ = m[0, 0] * x2 + m[0, 1] * y2 + m[0, 2]
        y2_ = m[1, 0] * x2 + m[1, 1] * y2 + m[1, 2]
        w, h = x2_ - x1_, y2_ - y1_
        cx, cy = x1_ + w / 2, y1_ + h / 2
        self.mean[:4] = [cx, cy, w / h, h]

    def update(self, detection, class_id, conf):
        self.conf = conf
        self.class_id = int(class_id)
        self.mean, self.covariance = self.kf.update(
            self.mean, self.covariance, detection.to_xyah(), detection.confidence)
        feature = detection.feature / _l2norm32(detection.feature)
        # python-float weights applied to float32 arrays: each weight is the
        # float64 value rounded once to float32 (NumPy weak-scalar rule)
        a = np.float32(self.ema_alpha)
        b = np.float32(1.0 - self.ema_alpha)
        smooth_feat = a * self.features[-1] + b * feature
        smooth_feat = smooth_feat / _l2norm32(smooth_feat)
        self.features = [smooth_feat.astype(np.float32)]
        self.hits += 1
        self.time_since_update = 0
        if self.state == TENTATIVE and self.hits >= self._n_init:
            self.state = CONFIRMED

    def mark_missed(self):
        if self.state == TENTATIVE:
            self.state = DELETED
        elif self.time_since_update > self._max_age:
            self.state = DELETED

    def is_tentative(self):
        return self.state == TENTATIVE

    def is_confirmed(self):
        return self.state == CONFIRMED

    def is_deleted(self):
        return self.state == DELETED


# ----------------------------------------------------------------------------
# A.5 appearance metric
# ----------------------------------------------------------------------------
def _cosine_distance(a, b):
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    a = a / np.linalg.norm(a, axis=1, keepdims=True)
    b = b / np.linalg.norm(b, axis=1, keepdims=True)
    return np.float32(1.0) - np.dot(a, b.T)


def _nn_cosine_distance(x, y):
    return _cosine_distance(x, y).min(axis=0)


class NearestNeighborDistanceMetric:
    def __init__(self, matching_threshold, budget=None):
        self.matching_threshold = matching_threshold
        self.budget = budget
        self.samples = {}

    def partial_fit(self, features, targets, active_targets):
        for feature, target in zip(features, targets):
            self.samples.setdefault(target, []).append(feature)
            if self.budget is not None:
                self.samples[target] = self.samples[target][-self.budget:]
        self.samples = {k: self.samples[k] for k in active_targets}

    def distance(self, features, targets):
        cost_matrix = np.zeros((len(targets), len(features)), dtype=np.float64)
        for i, target in enumerate(targets):
            cost_matrix[i, :] = _nn_cosine_distance(self.samples[target], features)
        return cost_matrix


# ----------------------------------------------------------------------------
# A.7 IoU cost
# ----------------------------------------------------------------------------
def iou(bbox, candidates):
    """bbox: float64 tlwh of one track; candidates: float32 tlwh [n,4].
    pin: candidate bottom-right and area are evaluated in float32 (array ops on
    float32), everything that mixes with the float64 track box is float64."""
    bbox = np.asarray(bbox, dtype=np.float64)
    candidates = np.asarray(candidates, dtype=np.float32)
    bbox_tl, bbox_br = bbox[:2], bbox[:2] + bbox[2:]
    candidates_tl = candidates[:, :2].astype(np.float64)
    candidates_br = (candidates[:, :2] + candidates[:, 2:]).astype(np.float64)
    tl = np.maximum(bbox_tl[None, :], candidates_tl)
    br = np.minimum(bbox_br[None, :], candidates_br)
    wh = np.maximum(0.0, br - tl)
    area_intersection = wh[:, 0] * wh[:, 1]
    area_bbox = bbox[2] * bbox[3]
    area_candidates = (candidates[:, 2] * candidates[:, 3]).astype(np.float64)
    return area_intersection / (area_bbox + area_candidates - area_intersection)


def iou_cost(tracks, detections, track_indices, detection_indices):
    cost_matrix = np.zeros((len(track_indices), len(detection_indices)),
                           dtype=np.float64)
    candidates = np.asarray([detections[i].tlwh for i in detection_indices],
                            dtype=np.float32).reshape(-1, 4)
    for row, track_idx in enumerate(track_indices):
        if tracks[track_idx].time_since_update > 1:
            cost_matrix[row, :] = INFTY_COST
            continue
        bbox = tracks[track_idx].to_tlwh()
        cost_matrix[row, :] = 1.0 - iou(bbox, candidates)
    return cost_matrix


# ----------------------------------------------------------------------------
# A.6 association
# ----------------------------------------------------------------------------
def min_cost_matching(distance_metric, max_distance, tracks, detections,
                      track_indices, detection_indices, trace=None, stage=""):
    if len(detection_indices) == 0 or len(track_indices) == 0:
        return [], list(track_indices), list(detection_indices)

    cost_matrix = distance_metric(tracks, detections, track_indices,
                                  detection_indices)
    raw = cost_matrix.copy() if trace is not None else None
    cost_matrix[cost_matrix > max_distance] = max_distance + 1e-5
    row_indices, col_indices = linear_sum_assignment(cost_matrix)
    if trace is not None:
        trace[stage + "_raw"] = raw
        trace[stage + "_cost"] = cost_matrix.copy()
        trace[stage + "_rows"] = np.asarray(row_indices).copy()
        trace[stage + "_cols"] = np.asarray(col_indices).copy()
        trace[stage + "_track_indices"] = np.asarray(track_indices, dtype=np.int64)
        trace[stage + "_detection_indices"] = np.asarray(detection_indices, dtype=np.int64)

    matches, unmatched_tracks, unmatched_detections = [], [], []
    col_set = set(int(c) for c in col_indices)
    row_set = set(int(r) for r in row_indices)
    for col, detection_idx in enumerate(detection_indices):
        if col not in col_set:
            unmatched_detections.append(detection_idx)
    for row, track_idx in enumerate(track_indices):
        if row not in row_set:
            unmatched_tracks.append(track_idx)
    for row, col in zip(row_indices, col_indices):
        track_idx = track_indices[row]
        detection_idx = detection_indices[col]
        if cost_matrix[row, col] > max_distance:
            unmatched_tracks.append(track_idx)
            unmatched_detections.append(detection_idx)
        else:
            matches.append((track_idx, detection_idx))
    return matches, unmatched_tracks, unmatched_detections


def gate_cost_matrix(kf, cost_matrix, tracks, detections, track_indices,
                     detection_indices, mc_lambda, gated_cost=INFTY_COST):
    gating_threshold = chi2inv95[4]
    measurements = np.asarray(
        [detections[i].to_xyah() for i in detection_indices], dtype=np.float32)
    for row, track_idx in enumerate(track_indices):
        track = tracks[track_idx]
        gating_distance = kf.gating_distance(track.mean, track.covariance,
                                             measurements)
        cost_matrix[row, gating_distance > gating_threshold] = gated_cost
        cost_matrix[row] = mc_lambda * cost_matrix[row] + \
            (1 - mc_lambda) * gating_distance
    return cost_matrix


class Tracker:
    def __init__(self, metric, max_iou_distance=0.7, max_age=30, n_init=3,
                 _lambda=0.0, ema_alpha=0.9, mc_lambda=0.995):
        self.metric = metric
        self.max_iou_distance = max_iou_distance
        self.max_age = max_age
        self.n_init = n_init
        self._lambda = _lambda
        self.ema_alpha = ema_alpha
        self.mc_lambda = mc_lambda
        self.kf = KalmanFilter()
        self.tracks = []
        self._next_id = 1
        self.trace = None  # dict filled per update() when tracing is on

    def predict(self):
        for track in self.tracks:
            track.predict()

    def increment_ages(self):
        """Upstream Tracker.increment_ages (its stream loop calls StrongSORT.increment_ages() on frames
        without detections, SURVEY A.2 caller side): age and time_since_update advance and every
        track is marked missed; no Kalman predict.  Deleted tracks stay listed until the next update()."""
        for track in self.tracks:
            track.age += 1
            track.time_since_update += 1
            track.mark_missed()

    def camera_update(self, warp_matrix):
        """One warp per frame applied to every track (upstream estimates the same ECC warp once
        per track)."""
        for track in self.tracks:
            track.camera_update(warp_matrix)

    def update(self, detections, classes, confidences):
        matches, unmatched_tracks, unmatched_detections = self._match(detections)
        if self.trace is not None:
            self.trace["matches"] = np.asarray(matches, dtype=np.int64).reshape(-1, 2)
            self.trace["unmatched_tracks"] = np.asarray(sorted(unmatched_tracks), dtype=np.int64)
            self.trace["unmatched_detections"] = np.asarray(unmatched_detections, dtype=np.int64)

        for track_idx, detection_idx in matches:
            self.tracks[track_idx].update(detections[detection_idx],
                                          classes[detection_idx],
                                          confidences[detection_idx])
        for track_idx in unmatched_tracks:
            self.tracks[track_idx].mark_missed()
        for detection_idx in unmatched_detections:
            self._initiate_track(detections[detection_idx],
                                 classes[detection_idx],
                                 confidences[detection_idx])
        self.tracks = [t for t in self.tracks if not t.is_deleted()]

        active_targets = [t.track_id for t in self.tracks if t.is_confirmed()]
        features, targets = [], []
        for track in self.tracks:
            if not track.is_confirmed():
                continue
            features += track.features
            targets += [track.track_id for _ in track.features]
        self.metric.partial_fit(features, targets, active_targets)

    def _match(self, detections):
        def gated_metric(tracks, dets, track_indices, detection_indices):
            features = np.array([dets[i].feature for i in detection_indices],
                                dtype=np.float32)
            targets = [tracks[i].track_id for i in track_indices]
            cost_matrix = self.metric.distance(features, targets)
            if self.trace is not None:
                self.trace["A_appearance"] = cost_matrix.copy()
            return gate_cost_matrix(self.kf, cost_matrix, tracks, dets,
                                    track_indices, detection_indices,
                                    self.mc_lambda)

        confirmed_tracks = [i for i, t in enumerate(self.tracks) if t.is_confirmed()]
        unconfirmed_tracks = [i for i, t in enumerate(self.tracks) if not t.is_confirmed()]

        # Stage A ("matching_cascade" of this fork: a single min_cost_matching
        # over all confirmed tracks; its unmatched-track list is rebuilt as
        # set(track_indices) - matched; pin: ascending order).
        detection_indices = list(range(len(detections)))
        matches_a, _, unmatched_detections = min_cost_matching(
            gated_metric, self.metric.matching_threshold, self.tracks,
            detections, confirmed_tracks, detection_indices,
            trace=self.trace, stage="A")
        matched_a = set(k for k, _ in matches_a)
        unmatched_tracks_a = [k for k in confirmed_tracks if k not in matched_a]

        # Stage B: IoU on unconfirmed + just-missed tracks
        iou_track_candidates = unconfirmed_tracks + [
            k for k in unmatched_tracks_a if self.tracks[k].time_since_update == 1]
        unmatched_tracks_a = [
            k for k in unmatched_tracks_a if self.tracks[k].time_since_update != 1]
        matches_b, unmatched_tracks_b, unmatched_detections = min_cost_matching(
            iou_cost, self.max_iou_distance, self.tracks, detections,
            iou_track_candidates, unmatched_detections,
            trace=self.trace, stage="B")

        matches = matches_a + matches_b
        unmatched_tracks = list(set(unmatched_tracks_a + unmatched_tracks_b))
        return matches, unmatched_tracks, unmatched_detections

    def _initiate_track(self, detection, class_id, conf):
        mean, covariance = self.kf.initiate(detection.to_xyah())
        self.tracks.append(Track(mean, covariance, self._next_id, int(class_id),
                                 conf, self.n_init, self.max_age,
                                 self.ema_alpha, detection.feature))
        self._next_id += 1


# ----------------------------------------------------------------------------
# A.2 / A.3 StrongSORT front object
# ----------------------------------------------------------------------------
def xyxy2xywh(x):
    x = np.asarray(x, dtype=np.float32)
    y = np.empty_like(x)
    y[:, 0] = (x[:, 0] + x[:, 2]) / np.float32(2)
    y[:, 1] = (x[:, 1] + x[:, 3]) / np.float32(2)
    y[:, 2] = x[:, 2] - x[:, 0]
    y[:, 3] = x[:, 3] - x[:, 1]
    return y


def xywh_to_tlwh(bbox_xywh):
    bbox_xywh = np.asarray(bbox_xywh, dtype=np.float32)
    t = bbox_xywh.copy()
    t[:, 0] = bbox_xywh[:, 0] - bbox_xywh[:, 2] / np.float32(2.0)
    t[:, 1] = bbox_xywh[:, 1] - bbox_xywh[:, 3] / np.float32(2.0)
    return t


def crop_box_xyxy(box_xywh, width, height):
    """A.3: int() truncation toward zero of float32 sums, clamped."""
    x, y, w, h = [np.float32(v) for v in box_xywh]
    x1 = max(int(x - w / np.float32(2)), 0)
    x2 = min(int(x + w / np.float32(2)), width - 1)
    y1 = max(int(y - h / np.float32(2)), 0)
    y2 = min(int(y + h / np.float32(2)), height - 1)
    return x1, y1, x2, y2


class StrongSORTOracle:
    """``StrongSORT.update(dets, img)`` restated.

    ``extractor(img, boxes_xyxy_int[N,4]) -> float32 [N,512]`` supplies the ReID
    embeddings (oracle/osnet_torch.py) or any stand-in for tracker-only tests.
    """

    def __init__(self, extractor, max_dist=0.2, max_iou_distance=0.7,
                 max_age=30, n_init=3, nn_budget=100, mc_lambda=0.995,
                 ema_alpha=0.9):
        self.extractor = extractor
        self.max_dist = max_dist
        metric = NearestNeighborDistanceMetric(self.max_dist, nn_budget)
        self.tracker = Tracker(metric, max_iou_distance=max_iou_distance,
                               max_age=max_age, n_init=n_init,
                               ema_alpha=ema_alpha, mc_lambda=mc_lambda)
        self.trace_enabled = False
        self.last_trace = None

    def update(self, dets, ori_img, features=None):
        dets = np.asarray(dets, dtype=np.float32).reshape(-1, 6)
        xyxys, confs, clss = dets[:, 0:4], dets[:, 4], dets[:, 5]
        xywhs = xyxy2xywh(xyxys)
        self.height, self.width = ori_img.shape[:2]

        if features is None:
            boxes = np.asarray([crop_box_xyxy(b, self.width, self.height)
                                for b in xywhs], dtype=np.int64).reshape(-1, 4)
            features = self.extractor(ori_img, boxes) if len(boxes) else \
                np.zeros((0, 512), dtype=np.float32)
        features = np.asarray(features, dtype=np.float32)
        bbox_tlwh = xywh_to_tlwh(xywhs)
        detections = [Detection(bbox_tlwh[i], confs[i], features[i])
                      for i in range(len(confs))]

        self.tracker.trace = {} if self.trace_enabled else None
        if self.trace_enabled:
            self.tracker.trace["features"] = features.copy()
        self.tracker.predict()
        if self.trace_enabled:
            self.tracker.trace["pred_mean"] = np.asarray(
                [t.mean for t in self.tracker.tracks], dtype=np.float64).reshape(-1, 8)
            self.tracker.trace["pred_cov"] = np.asarray(
                [t.covariance for t in self.tracker.tracks], dtype=np.float64).reshape(-1, 8, 8)
        self.tracker.update(detections, clss, confs)
        self.last_trace = self.tracker.trace

        outputs = []
        for track in self.tracker.tracks:
            if not track.is_confirmed() or track.time_since_update > 1:
                continue
            x, y, w, h = track.to_tlwh()
            x1 = max(int(x), 0)
            x2 = min(int(x + w), self.width - 1)
            y1 = max(int(y), 0)
            y2 = min(int(y + h), self.height - 1)
            outputs.append(np.array([x1, y1, x2, y2, track.track_id,
                                     track.class_id, track.conf],
                                    dtype=np.float64))
        if len(outputs) > 0:
            return np.stack(outputs, axis=0)
        return np.zeros((0, 7), dtype=np.float64)

    def increment_ages(self):
        self.tracker.increment_ages()

    # introspection used by the parity tests ------------------------------
    def track_table(self):
        t = self.tracker.tracks
        return {
            "track_id": np.asarray([x.track_id for x in t], dtype=np.int64),
            "state": np.asarray([x.state for x in t], dtype=np.int64),
            "hits": np.asarray([x.hits for x in t], dtype=np.int64),
            "age": np.asarray([x.age for x in t], dtype=np.int64),
            "tsu": np.asarray([x.time_since_update for x in t], dtype=np.int64),
            "mean": np.asarray([x.mean for x in t], dtype=np.float64).reshape(-1, 8),
            "cov": np.asarray([x.covariance for x in t], dtype=np.float64).reshape(-1, 8, 8),
            "feat": np.asarray([x.features[-1] for x in t], dtype=np.float32).reshape(-1, 512)
            if len(t) and len(t[0].features[-1]) == 512 else
            np.asarray([x.features[-1] for x in t], dtype=np.float32),
            "gallery_len": np.asarray(
                [len(self.tracker.metric.samples.get(x.track_id, [])) for x in t],
                dtype=np.int64),
        }
