"""Generate the committed golden fixtures under tests/golden/ from the CPU
oracle (self-pinned: the reference ships no fixtures).

  c1_e2e.npz      C1 (640x640, 10 objects, 8 frames): dets, oracle OSNet
                  embeddings, StrongSORT outputs per frame, final track table
  c2_tracker.npz  C2-sized tracker-only run (100 dets/frame, 60 frames) on
                  FeatureBank embeddings: outputs per frame + decision margins
  reid_kat.npz    one 320x320 frame, 6 crop boxes, oracle embeddings
  nms_kat.npz     a decoded YOLOv8 head [84,1200] and torchvision's NMS result
  decode_kat.npz  raw YOLOv8 detect and pose heads (96x160 input, A = 315) and their decode
  camera_kat.npz  a tracker state before / after Tracker.camera_update with a Euclidean warp
  gallery_kat.npz 3 streams' exported tracks and the cross-stream match of every stream

Run from the repo root:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import gallery_np, nms_np, osnet_torch, strongsort_np as ss, yolo_decode_np  # noqa: E402
from strongsort_yolo_b200 import synth, weights, yolo  # noqa: E402
from helpers import FeatureBank  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def pack_rows(rows_per_frame):
    lens = np.asarray([len(r) for r in rows_per_frame], dtype=np.int64)
    flat = np.concatenate([np.asarray(r, dtype=np.float64).reshape(-1, 7) for r in rows_per_frame], 0)
    return flat, lens


def c1_e2e():
    ext = osnet_torch.OracleExtractor(weights.load_state_dict())
    st = synth.make_stream("C1")
    trk = ss.StrongSORTOracle(ext)
    trk.trace_enabled = True
    dets, feats, rows = [], [], []
    for _ in range(8):
        fr = st.next_frame()
        out = trk.update(fr.dets, fr.img)
        dets.append(fr.dets)
        feats.append(trk.last_trace["features"])
        rows.append(out)
    flat, lens = pack_rows(rows)
    tab = trk.track_table()
    np.savez_compressed(
        os.path.join(OUT, "c1_e2e.npz"), rows=flat, row_lens=lens,
        det_lens=np.asarray([len(d) for d in dets]), dets=np.concatenate(dets, 0),
        feats=np.concatenate(feats, 0).astype(np.float32),
        **{"tab_" + k: v for k, v in tab.items()}, next_id=trk.tracker._next_id)
    print("c1_e2e: rows/frame", lens.tolist(), "next_id", trk.tracker._next_id)


def c2_tracker():
    st = synth.make_stream("C2", render=False)
    bank = FeatureBank(seed=7)
    trk = ss.StrongSORTOracle(None)
    trk.trace_enabled = True
    rows, margins = [], []
    img = np.zeros((1080, 1920, 3), dtype=np.uint8)
    for _ in range(60):
        fr = st.next_frame()
        out = trk.update(fr.dets, img, features=bank(fr.gt_ids))
        rows.append(out)
        tr = trk.last_trace
        m = np.inf
        for stg, thr in (("A", 0.2), ("B", 0.7)):
            if stg + "_raw" in tr and tr[stg + "_raw"].size:
                raw = tr[stg + "_raw"]
                m = min(m, np.abs(raw[raw < 1e4] - thr).min() if (raw < 1e4).any() else np.inf)
        margins.append(m)
    flat, lens = pack_rows(rows)
    np.savez_compressed(os.path.join(OUT, "c2_tracker.npz"), rows=flat, row_lens=lens,
                        margins=np.asarray(margins), next_id=trk.tracker._next_id)
    print("c2_tracker: frames", len(lens), "rows last", lens[-1], "next_id",
          trk.tracker._next_id, "min threshold margin", np.min(margins))


def reid_kat():
    ext = osnet_torch.OracleExtractor(weights.load_state_dict())
    st = synth.SyntheticStream(width=320, height=320, n_persistent=6, seed=4242)
    fr = st.next_frame()
    xywh = ss.xyxy2xywh(fr.dets[:, :4])
    boxes = np.asarray([ss.crop_box_xyxy(b, 320, 320) for b in xywh], dtype=np.int32)
    # plus crops larger than 256x128 (down-sampling) and a 1-pixel-wide sliver
    boxes = np.concatenate([boxes, np.asarray([[10, 8, 300, 312], [0, 0, 319, 319],
                                               [100, 20, 101, 300]], dtype=np.int32)], 0)
    emb = ext(fr.img, boxes)
    np.savez_compressed(os.path.join(OUT, "reid_kat.npz"), img=fr.img, dets=fr.dets, boxes=boxes, emb=emb)
    print("reid_kat: boxes", boxes.tolist(), "emb norm", np.linalg.norm(emb, axis=1))


def nms_kat():
    rng = np.random.default_rng(99)
    dets = np.zeros((40, 6), dtype=np.float32)
    cx, cy = rng.uniform(60, 580, 40), rng.uniform(60, 580, 40)
    w, h = rng.uniform(20, 90, 40), rng.uniform(30, 120, 40)
    dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3] = cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2
    dets[:, 4] = rng.uniform(0.35, 0.95, 40)
    dets[:, 5] = rng.integers(0, 3, 40)
    pred = yolo.synth_head(dets, num_classes=80, num_anchors=1200, rng=rng, jitter=4)
    out = nms_np.yolo_nms(pred, 80, 0, 0.3, 0.4, 1000, False)
    np.savez_compressed(os.path.join(OUT, "nms_kat.npz"), pred=pred, out=out)
    print("nms_kat: kept", len(out))


def decode_kat():
    rng = np.random.default_rng(314)
    in_h, in_w = 96, 160                                   # A = 12*20 + 6*10 + 3*5 = 315
    raw_det = rng.normal(0, 2.0, (64 + 5, 315)).astype(np.float32)
    raw_pose = rng.normal(0, 2.0, (64 + 1 + 3 * 17, 315)).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "decode_kat.npz"), in_hw=np.asarray([in_h, in_w]),
                        raw_det=raw_det, out_det=yolo_decode_np.decode_v8(raw_det, 5, 0, in_h, in_w),
                        raw_pose=raw_pose, out_pose=yolo_decode_np.decode_v8(raw_pose, 1, 17, in_h, in_w))
    print("decode_kat: 315 anchors, det + pose heads")


def camera_kat():
    st = synth.make_stream("C1", render=False)
    bank = FeatureBank(seed=11)
    trk = ss.StrongSORTOracle(None)
    img = np.zeros((640, 640, 3), dtype=np.uint8)
    frames = []
    for _ in range(5):
        fr = st.next_frame()
        frames.append(fr)
        trk.update(fr.dets, img, features=bank(fr.gt_ids))
    th = np.deg2rad(0.7)
    warp = np.array([[np.cos(th), -np.sin(th), 4.5], [np.sin(th), np.cos(th), -2.25]])
    before = np.stack([t.mean.copy() for t in trk.tracker.tracks])
    trk.tracker.camera_update(warp)
    after = np.stack([t.mean.copy() for t in trk.tracker.tracks])
    np.savez_compressed(os.path.join(OUT, "camera_kat.npz"), warp=warp, mean_before=before, mean_after=after)
    print("camera_kat:", len(before), "tracks")


def gallery_kat():
    rng = np.random.default_rng(2718)
    G, T, D = 3, 16, 512
    base = np.maximum(rng.normal(0, 1, (14, D)), 0)
    feat = np.zeros((G, T, D), dtype=np.float32)
    ids = np.full((G, T), -1, dtype=np.int32)
    for g in range(G):
        n = 8 + 2 * g
        who = rng.permutation(14)[:n]
        f = base[who] + rng.normal(0, 0.04, (n, D))
        feat[g, :n] = (f / np.linalg.norm(f, axis=1, keepdims=True)).astype(np.float32)
        ids[g, :n] = 1 + np.arange(n) + 50 * g
    res = [gallery_np.cross_match(feat[g], ids[g], feat, ids, g, 0.2) for g in range(G)]
    np.savez_compressed(os.path.join(OUT, "gallery_kat.npz"), feat=feat, ids=ids,
                        m_rank=np.stack([r[0] for r in res]), m_id=np.stack([r[1] for r in res]),
                        m_dist=np.stack([r[2] for r in res]))
    print("gallery_kat: matches per stream", [(r[0] >= 0).sum() for r in res])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--new-only" not in sys.argv:
        c1_e2e(); c2_tracker(); reid_kat(); nms_kat()
    decode_kat(); camera_kat(); gallery_kat()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
