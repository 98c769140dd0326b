"""CPU suite: the oracle against its own committed golden vectors (self-pinned;
the reference ships none), plus oracle self-consistency checks."""
import os

import numpy as np
import pytest

from helpers import FeatureBank
from oracle import nms_np, osnet_torch, strongsort_np as ss
from strongsort_yolo_b200 import synth


def _unpack(flat, lens):
    out, o = [], 0
    for n in lens:
        out.append(flat[o:o + n]); o += n
    return out


def test_osnet_mac_count():
    # SURVEY.md Appendix B self-check: 82.3 MMAC per 3x256x128 crop
    assert osnet_torch.count_macs() == 82314880


def test_reid_kat(oracle_extractor, golden_dir):
    g = np.load(os.path.join(golden_dir, "reid_kat.npz"))
    emb = oracle_extractor(g["img"], g["boxes"])
    np.testing.assert_allclose(emb, g["emb"], rtol=1e-4, atol=1e-4)


def test_c1_e2e_golden(oracle_extractor, golden_dir):
    g = np.load(os.path.join(golden_dir, "c1_e2e.npz"))
    st = synth.make_stream("C1")
    trk = ss.StrongSORTOracle(oracle_extractor)
    rows = _unpack(g["rows"], g["row_lens"])
    dets = _unpack(g["dets"], g["det_lens"])
    for f in range(8):
        fr = st.next_frame()
        np.testing.assert_array_equal(fr.dets, dets[f])      # the scene generator is pinned too
        out = trk.update(fr.dets, fr.img)
        np.testing.assert_array_equal(out[:, :6], rows[f][:, :6])
        np.testing.assert_allclose(out[:, 6], rows[f][:, 6], atol=1e-7)
    assert trk.tracker._next_id == int(g["next_id"])
    tab = trk.track_table()
    np.testing.assert_array_equal(tab["track_id"], g["tab_track_id"])
    np.testing.assert_allclose(tab["mean"], g["tab_mean"], rtol=1e-9)


def test_c2_tracker_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "c2_tracker.npz"))
    st = synth.make_stream("C2", render=False)
    bank = FeatureBank(seed=7)
    trk = ss.StrongSORTOracle(None)
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    rows = _unpack(g["rows"], g["row_lens"])
    for f in range(len(rows)):
        fr = st.next_frame()
        out = trk.update(fr.dets, img, features=bank(fr.gt_ids))
        np.testing.assert_array_equal(out[:, :6], rows[f][:, :6])
    assert trk.tracker._next_id == int(g["next_id"])


def test_nms_kat(golden_dir):
    g = np.load(os.path.join(golden_dir, "nms_kat.npz"))
    out = nms_np.yolo_nms(g["pred"], 80, 0, 0.3, 0.4, 1000, False)
    np.testing.assert_array_equal(out, g["out"])


def test_kf_gating_matches_direct_inverse():
    kf = ss.KalmanFilter()
    rng = np.random.default_rng(3)
    m, P = kf.initiate(np.array([300., 200., 0.4, 120.], dtype=np.float32))
    for _ in range(3):
        m, P = kf.predict(m, P)
    z = np.array([[301., 203., 0.41, 118.], [500., 100., 0.3, 90.]], dtype=np.float32)
    g = kf.gating_distance(m, P, z)
    mu, S = kf.project(m, P)
    d = z.astype(np.float64) - mu
    ref = np.einsum("ni,ij,nj->n", d, np.linalg.inv(S), d)
    np.testing.assert_allclose(g, ref, rtol=1e-9)


def test_empty_and_ragged_frames():
    trk = ss.StrongSORTOracle(None)
    img = np.zeros((480, 640, 3), dtype=np.uint8)
    out = trk.update(np.zeros((0, 6), np.float32), img, features=np.zeros((0, 512), np.float32))
    assert out.shape == (0, 7)
    bank = FeatureBank(seed=1)
    st = synth.SyntheticStream(width=640, height=480, n_persistent=5, seed=5, render=False)
    for f in range(6):
        fr = st.next_frame()
        k = [5, 0, 3, 5, 1, 5][f]
        out = trk.update(fr.dets[:k], img, features=bank(fr.gt_ids[:k]))
    assert out.shape[1] == 7


def test_decode_golden(golden_dir):
    from oracle import yolo_decode_np
    g = np.load(os.path.join(golden_dir, "decode_kat.npz"))
    in_h, in_w = [int(v) for v in g["in_hw"]]
    np.testing.assert_allclose(yolo_decode_np.decode_v8(g["raw_det"], 5, 0, in_h, in_w), g["out_det"], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(yolo_decode_np.decode_v8(g["raw_pose"], 1, 17, in_h, in_w), g["out_pose"], rtol=1e-6, atol=1e-5)


def test_camera_update_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "camera_kat.npz"))
    for before, after in zip(g["mean_before"], g["mean_after"]):
        t = ss.Track(before.copy(), np.eye(8), 1, 0, 0.5, 3, 30, 0.9, None)
        t.camera_update(g["warp"])
        np.testing.assert_allclose(t.mean, after, rtol=1e-13, atol=0)


def test_gallery_cross_match_golden(golden_dir):
    from oracle import gallery_np
    g = np.load(os.path.join(golden_dir, "gallery_kat.npz"))
    for r in range(g["feat"].shape[0]):
        m_rank, m_id, m_dist = gallery_np.cross_match(g["feat"][r], g["ids"][r], g["feat"], g["ids"], r, 0.2)
        np.testing.assert_array_equal(m_rank, g["m_rank"][r])
        np.testing.assert_array_equal(m_id, g["m_id"][r])
        live = g["ids"][r] >= 0
        np.testing.assert_allclose(m_dist[live], g["m_dist"][r][live], atol=1e-6)
