"""GPU: YOLOv8 head decode (csrc/decode.cu) and the batched camera-motion update against the
CPU restatements, through the C-ABI."""
import numpy as np
import pytest

from helpers import FeatureBank, assert_rows_equal, assert_tables_equal
from oracle import nms_np, strongsort_np as ss, yolo_decode_np
from strongsort_yolo_b200 import synth, yolo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("in_h,in_w,nc,nk", [(640, 640, 80, 0), (384, 640, 80, 0), (640, 640, 1, 17)])
def test_decode_v8_matches_restatement(in_h, in_w, nc, nk):
    import torch
    rng = np.random.default_rng(in_h + nk)
    dec = yolo.YoloV8Decode(nc, nk, in_h, in_w)
    raw = rng.normal(0, 2.0, (64 + nc + 3 * nk, dec.A)).astype(np.float32)
    got = dec(torch.as_tensor(raw).cuda()).cpu().numpy()
    want = yolo_decode_np.decode_v8(raw, nc, nk, in_h, in_w)
    # float32 expf / division differ by a few ulp between libm and CUDA: 1e-5 relative on boxes
    # (pixels up to ~1e3), 1e-6 absolute on probabilities
    np.testing.assert_allclose(got[:4], want[:4], rtol=2e-5, atol=2e-3)
    np.testing.assert_allclose(got[4:4 + nc], want[4:4 + nc], rtol=1e-5, atol=1e-6)
    if nk:
        k_got, k_want = got[4 + nc:].reshape(nk, 3, -1), want[4 + nc:].reshape(nk, 3, -1)
        np.testing.assert_allclose(k_got[:, :2], k_want[:, :2], rtol=1e-5, atol=1e-3)
        np.testing.assert_allclose(k_got[:, 2], k_want[:, 2], rtol=1e-5, atol=1e-6)


def test_decode_then_nms_recovers_detections():
    import torch
    rng = np.random.default_rng(11)
    in_h, in_w, nc = 384, 640, 80
    st = synth.make_stream("C1", render=False)
    d = st.next_frame().dets.copy()
    d[:, [0, 2]] *= in_w / 640.0; d[:, [1, 3]] *= in_h / 640.0          # into the network-input frame
    raw = yolo.synth_raw_head_v8(d, nc, in_h, in_w, rng=rng)
    dec = yolo.YoloV8Decode(nc, 0, in_h, in_w)
    nms = yolo.YoloNMS(num_classes=nc, max_anchors=dec.A)
    rows = nms.detect(dec(torch.as_tensor(raw).cuda()))
    want = nms_np.yolo_nms(yolo_decode_np.decode_v8(raw, nc, 0, in_h, in_w), nc, 0, 0.3, 0.4, 1000, False)
    assert rows.shape == want.shape
    np.testing.assert_allclose(rows[:, :4], want[:, :4], rtol=2e-5, atol=2e-3)
    np.testing.assert_allclose(rows[:, 4], want[:, 4], atol=1e-6)
    np.testing.assert_array_equal(rows[:, 5], want[:, 5])


def test_c2_head_decode_nms_scale_recovers_frame_detections():
    """The bench's detector post-process chain on a 1080p frame letterboxed to 384x640: raw v8 head ->
    decode -> NMS -> scale_boxes gives the frame's detections back (bit-exact vs the CPU restatement of
    scale_boxes; within the DFL round trip of the detections that generated the head)."""
    import torch
    rng = np.random.default_rng(3)
    net, frame = (384, 640), (1080, 1920)
    fr = synth.make_stream("C2", render=False).next_frame()
    g, px, py = yolo.letterbox_params(net, frame)
    fr.dets[:, 5] = np.where(fr.gt_ids >= 0, fr.gt_ids % 80, 79)       # one class per identity: class-aware NMS keeps overlaps
    d = fr.dets.copy()
    d[:, [0, 2]] = d[:, [0, 2]] * g + px
    d[:, [1, 3]] = d[:, [1, 3]] * g + py
    raw = yolo.synth_raw_head_v8(d, 80, net[0], net[1], rng=rng)
    dec = yolo.YoloV8Decode(80, 0, net[0], net[1])
    nms = yolo.YoloNMS(num_classes=80, max_anchors=dec.A)
    nms(dec(torch.as_tensor(raw).cuda()))
    unscaled = nms._out.clone()
    out, cnt = nms.scale_boxes(net, frame)
    m = int(cnt[0].item())
    got = out[:m].cpu().numpy()
    want = nms_np.scale_boxes(unscaled[:m].cpu().numpy(), net, frame)
    np.testing.assert_array_equal(got, want)
    assert m == len(fr.dets)
    order = np.argsort(-fr.dets[:, 4], kind="stable")
    np.testing.assert_allclose(got[:, :4], fr.dets[order, :4], rtol=0, atol=0.05)
    np.testing.assert_allclose(got[:, 4], fr.dets[order, 4], rtol=0, atol=1e-5)
    np.testing.assert_array_equal(got[:, 5], fr.dets[order, 5])


def test_fused_postprocess_equals_the_three_calls():
    """ssb_yolo_postprocess_v8 (one call, 5 launches, look-ahead tickets) == decode -> NMS -> scale_boxes, bit for bit."""
    import torch
    rng = np.random.default_rng(5)
    net, frame = (384, 640), (1080, 1920)
    st = synth.make_stream("C2", render=False)
    dec = yolo.YoloV8Decode(80, 0, net[0], net[1])
    nms = yolo.YoloNMS(num_classes=80, max_anchors=dec.A)
    post = yolo.YoloV8Post(80, 0, net[0], net[1], frame, depth=3)
    g, px, py = yolo.letterbox_params(net, frame)
    raws, want = [], []
    for k in range(4):
        fr = st.next_frame()
        d = fr.dets.copy()
        d[:, 5] = np.where(fr.gt_ids >= 0, fr.gt_ids % 80, 79)
        d[:, [0, 2]] = d[:, [0, 2]] * g + px
        d[:, [1, 3]] = d[:, [1, 3]] * g + py
        raw = torch.as_tensor(yolo.synth_raw_head_v8(d, 80, net[0], net[1], rng=rng)).cuda()
        raws.append(raw)
        nms(dec(raw))
        out, cnt = nms.scale_boxes(net, frame)
        want.append(out[:int(cnt[0].item())].clone())
    tickets = [post.submit(r) for r in raws[:3]]            # three frames in flight
    got = [post.result(t).clone() for t in tickets] + [post(raws[3]).clone()]
    for a, b in zip(got, want):
        assert a.shape == b.shape and torch.equal(a, b)


def test_camera_update_matches_oracle_and_tracking_continues():
    from strongsort_yolo_b200.strong_sort import StrongSORT
    st = synth.make_stream("C1", render=False)
    bank = FeatureBank(seed=2)
    ora = ss.StrongSORTOracle(None)
    gpu = StrongSORT(max_tracks=128, max_dets=64)
    img = np.zeros((st.H, st.W, 3), dtype=np.uint8)
    th = np.deg2rad(0.4)
    warp = np.array([[np.cos(th), -np.sin(th), 3.25], [np.sin(th), np.cos(th), -1.5]])
    for f in range(12):
        fr = st.next_frame()
        feats = bank(fr.gt_ids)
        if f in (4, 5, 9):                 # the caller warps the tracks between frames (ECC on)
            ora.tracker.camera_update(warp)
            gpu.camera_update(None, None, warp_matrix=warp)
            assert_tables_equal(gpu.export_tracks(), ora.track_table(), rtol=1e-12)
        assert_rows_equal(gpu.update(fr.dets, img, features=feats), ora.update(fr.dets, img, features=feats))
    assert_tables_equal(gpu.export_tracks(), ora.track_table())


def test_c5_pose_pipeline_keypoints_follow_their_tracks():
    """Config C5's per-stream path: raw YOLOv8-pose head -> decode -> NMS (51 keypoint channels ride
    along) -> StrongSORT.update -> Results whose keypoints are re-indexed by the tracker's source
    detection index, consumed the way the reference's process() does (:45-62)."""
    import torch
    from strongsort_yolo_b200.results import results_from_tracks
    from strongsort_yolo_b200.strong_sort import StrongSORT
    in_h = in_w = 640
    st = synth.make_stream("C1", render=False)
    bank = FeatureBank(seed=6)
    dec = yolo.YoloV8Decode(1, 17, in_h, in_w)
    nms = yolo.YoloNMS(num_classes=1, num_extra=51, max_anchors=dec.A)
    trk = StrongSORT(max_tracks=128, max_dets=64)
    ora = ss.StrongSORTOracle(None)
    img = np.zeros((in_h, in_w, 3), dtype=np.uint8)
    rng = np.random.default_rng(3)
    seen_ids = 0
    for f in range(6):
        fr = st.next_frame()
        d = fr.dets.copy(); d[:, 5] = 0
        # keypoints = 17 points inside each box, derived from the box so they can be checked afterwards
        t = np.linspace(0.1, 0.9, 17)
        kp = np.stack([d[:, 0:1] + t * (d[:, 2:3] - d[:, 0:1]), d[:, 1:2] + t * (d[:, 3:4] - d[:, 1:2]),
                       np.ones((len(d), 17))], 2)
        raw = yolo.synth_raw_head_v8(d, 1, in_h, in_w, rng=rng, kpts=kp)
        rows = nms.detect(dec(torch.as_tensor(raw).cuda()))          # [M, 6 + 51]
        want = nms_np.yolo_nms(yolo_decode_np.decode_v8(raw, 1, 17, in_h, in_w), 1, 51, 0.3, 0.4, 1000, False)
        assert rows.shape == want.shape and len(rows) > 0
        # the detections feed tracker and oracle alike (embeddings from the bank, matched by position)
        order = np.argsort(-d[:, 4], kind="stable")
        feats = bank(fr.gt_ids[order][:len(rows)])
        out = trk.update(rows[:, :6], img, features=feats)
        ref = ora.update(rows[:, :6], img, features=feats)        # same detections for both trackers
        assert_rows_equal(out, ref)
        res = results_from_tracks(out, trk.last_det_index, keypoints=rows[:, 6:].reshape(-1, 17, 3))
        if res.boxes.id is None or len(res.boxes) == 0:
            continue
        ids = [int(b.id) for b in res.boxes]                          # the reference's :46
        assert len(set(ids)) == len(ids)
        seen_ids += len(ids)
        for k, (b, kpts) in enumerate(zip(res.boxes, res.keypoints)):  # the reference's :58-62
            di = int(trk.last_det_index[k])
            if di < 0:
                continue
            xy = np.asarray(kpts.xy.tolist())[0]
            np.testing.assert_allclose(xy, rows[di, 6:].reshape(17, 3)[:, :2], atol=1e-4)
            x1, y1, x2, y2 = rows[di, :4]
            assert (xy[:, 0] >= x1 - 1).all() and (xy[:, 0] <= x2 + 1).all()
            assert (xy[:, 1] >= y1 - 1).all() and (xy[:, 1] <= y2 + 1).all()
    assert seen_ids > 0


def test_decode_and_camera_against_committed_golden(golden_dir):
    """The committed fixtures (tools/make_golden.py): decode of raw detect / pose heads, camera warp."""
    import os
    import torch
    from strongsort_yolo_b200.strong_sort import StrongSORT
    g = np.load(os.path.join(golden_dir, "decode_kat.npz"))
    in_h, in_w = [int(v) for v in g["in_hw"]]
    for raw, out, nc, nk in ((g["raw_det"], g["out_det"], 5, 0), (g["raw_pose"], g["out_pose"], 1, 17)):
        got = yolo.YoloV8Decode(nc, nk, in_h, in_w)(torch.as_tensor(raw).cuda()).cpu().numpy()
        np.testing.assert_allclose(got[:4], out[:4], rtol=2e-5, atol=2e-3)
        np.testing.assert_allclose(got[4:4 + nc], out[4:4 + nc], rtol=1e-5, atol=1e-6)
        if nk:
            kg, ko = got[4 + nc:].reshape(nk, 3, -1), out[4 + nc:].reshape(nk, 3, -1)
            np.testing.assert_allclose(kg[:, :2], ko[:, :2], rtol=1e-5, atol=1e-3)
            np.testing.assert_allclose(kg[:, 2], ko[:, 2], rtol=1e-5, atol=1e-6)
    c = np.load(os.path.join(golden_dir, "camera_kat.npz"))
    # same stream / bank as tools/make_golden.py: camera_kat -> the table before the warp is the golden one
    st = synth.make_stream("C1", render=False)
    bank = FeatureBank(seed=11)
    gpu = StrongSORT(max_tracks=128, max_dets=64)
    img = np.zeros((640, 640, 3), dtype=np.uint8)
    for _ in range(5):
        fr = st.next_frame()
        gpu.update(fr.dets, img, features=bank(fr.gt_ids))
    np.testing.assert_allclose(gpu.export_tracks()["mean"], c["mean_before"], rtol=1e-9, atol=1e-9)
    gpu.camera_update(None, None, warp_matrix=c["warp"])
    np.testing.assert_allclose(gpu.export_tracks()["mean"], c["mean_after"], rtol=1e-9, atol=1e-9)


def test_decode_v5_matches_restatement_and_feeds_nms():
    """BASELINE config C1's detector head format (YOLOv5n: [1, 25200, 85] at 640x640)."""
    import torch
    rng = np.random.default_rng(21)
    nc = 80
    dec = yolo.YoloV5Decode(nc, 640, 640, conf=0.3)
    assert dec.A == 25200
    raw = rng.normal(-1.0, 2.0, (dec.A, 5 + nc)).astype(np.float32)
    got = dec(torch.as_tensor(raw).cuda()).cpu().numpy()
    want = yolo_decode_np.decode_v5(raw, nc, 640, 640, 0.3)
    np.testing.assert_allclose(got[:4], want[:4], rtol=2e-5, atol=2e-3)
    np.testing.assert_allclose(got[4:], want[4:], rtol=1e-5, atol=1e-6)
    assert ((got[4:] == 0) == (want[4:] == 0)).all()
    raw[:, 4] -= 6.0                                   # few candidates, like a real frame: run the NMS on them
    nms = yolo.YoloNMS(num_classes=nc, max_anchors=dec.A, conf=0.05)
    dec2 = yolo.YoloV5Decode(nc, 640, 640, conf=0.05)
    rows = nms.detect(dec2(torch.as_tensor(raw).cuda()))
    ref = nms_np.yolo_nms(yolo_decode_np.decode_v5(raw, nc, 640, 640, 0.05), nc, 0, 0.05, 0.4, 1000, False)
    assert rows.shape == ref.shape and len(rows) > 0
    np.testing.assert_allclose(rows[:, :5], ref[:, :5], rtol=2e-5, atol=2e-3)
    np.testing.assert_array_equal(rows[:, 5], ref[:, 5])
