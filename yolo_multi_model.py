#!/usr/bin/env python
"""CLI with the reference's surface (/root/reference/yolo_multi_model.py:341-354):

    python yolo_multi_model.py --source a.mp4 b.mp4 --track [--count]

``--source`` nargs+ (default '0'), ``--track``, ``--count``; one worker process
per source (the reference's ``multiprocessing.Pool``), stream *i* on GPU
*i mod G* instead of the hard-coded ``device=0`` (:41).  Per frame the worker
calls the B200 StrongSORT path where the reference calls ``model.track`` (:41),
builds the Results object ``process()`` consumes (:45-162) and appends the
reference's label lines (:165-169) to ``output/<name>_labels.txt``.

What is NOT here (out of scope): the YOLO backbone (no
weights offline), cv2 drawing / imshow / VideoWriter.  Detections therefore
come from one of:
  * ``synthetic[:C1|C2|C4[:frames]]``  -- the seeded synthetic stream; its
    detections go through a synthetic decoded head + the GPU NMS (yolo.py);
  * ``<video>`` with a MOT-style ``<video>.dets.txt`` sidecar
    (``frame,x1,y1,x2,y2,conf,cls`` per line), frames decoded with cv2.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from multiprocessing import get_context

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _frames(source):
    """yield (img uint8 HxWx3, dets float32 [N,6]) for one source."""
    if source.startswith("synthetic"):
        from strongsort_yolo_b200 import synth
        parts = source.split(":")
        cfg = parts[1] if len(parts) > 1 and parts[1] else "C1"
        n = int(parts[2]) if len(parts) > 2 else 30
        st = synth.make_stream(cfg)
        for _ in range(n):
            fr = st.next_frame()
            yield fr.img, fr.dets
        return
    import cv2
    side = source + ".dets.txt"
    if not os.path.exists(side):
        raise FileNotFoundError(f"{side}: detections sidecar required (no detector weights offline)")
    table = np.loadtxt(side, delimiter=",", ndmin=2).astype(np.float32)
    cap = cv2.VideoCapture(int(source) if source == "0" else source)
    if not cap.isOpened():
        print(f"Error: Could not open video file {source}.")        # reference :262-264
        return
    f = 0
    while True:
        f += 1
        ret, frame = cap.read()
        if not ret:
            break
        yield frame, table[table[:, 0] == f][:, 1:7]
    cap.release()


def process_video(args):
    """One worker = one source = one GPU stream (reference :244-339 without the cv2 drawing / imshow /
    VideoWriter).  The loop is a stream: frame k's host->device copy, detector
    post-process and OSNet run while frame k-1 is being associated (``update_pipelined``), the labels
    file is opened ONCE (the reference reopens it per frame, :39), and ``--count`` is a reduction over
    the device track table (``class_counts``) instead of re-parsing the whole labels file with pandas
    every frame (:288, O(frames^2))."""
    print(args)                                                       # reference :245
    source, track_, count_, gpu = args["source"], args["track"], args["count"], args["gpu"]
    import torch
    from strongsort_yolo_b200 import yolo
    from strongsort_yolo_b200.results import Boxes, Results, label_lines, results_from_tracks
    from strongsort_yolo_b200.strong_sort import StrongSORT
    dev = f"cuda:{gpu}"
    torch.cuda.set_device(gpu)
    name = os.path.splitext(os.path.basename(source.replace(":", "_")))[0]
    os.makedirs("output", exist_ok=True)
    labels_path = os.path.abspath(f"./output/{name}_labels.txt")
    if not track_ and count_:
        print("[INFO] count works only when objects are tracking.. so use: --track --count")   # :281
        return
    tracker = StrongSORT(device=dev) if track_ else None
    nms = yolo.YoloNMS(num_classes=80, max_anchors=8400, device=dev)
    rng = np.random.default_rng(0)
    t0, n_frames, n_rows = time.time(), 0, 0
    pending = []                 # (rows, det_index, shape) of frames whose results arrive one call later
    pin = None

    def emit(labels, rows, det_index, shape):
        nonlocal n_rows
        res = results_from_tracks(rows, det_index, orig_shape=shape)
        lines = label_lines(res)
        n_rows += len(lines)
        labels.writelines(lines)

    with open(labels_path, "a") as labels:                            # opened once, not per frame
        for img, dets in _frames(source):
            n_frames += 1
            # detector post-process on the GPU: decoded head -> conf filter -> class-aware NMS
            head = yolo.synth_head(dets, num_classes=80, num_anchors=8400, rng=rng)
            det = nms.detect(torch.as_tensor(head).to(dev))
            if not track_:
                res = Results(Boxes(det[:, :4], det[:, 4], det[:, 5], None), {0: "person"})
                labels.writelines(label_lines(res))
            elif len(det) == 0:
                # upstream's stream loop: no detections -> increment_ages() instead of update()
                rows = tracker.flush_pipelined()
                if rows is not None and pending:
                    emit(labels, rows, tracker.last_det_index, pending.pop(0))
                tracker.increment_ages()
            else:
                if pin is None or tuple(pin[0].shape) != tuple(img.shape):
                    pin = [torch.empty(img.shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
                slot = pin[n_frames & 1]                              # the copy of frame k-1 may still be in flight
                slot.copy_(torch.from_numpy(img))
                rows = tracker.update_pipelined(det[:, :6], slot)     # rows of the PREVIOUS frame
                pending.append(img.shape[:2])
                if rows is not None:
                    emit(labels, rows, tracker.last_det_index, pending.pop(0))
            if n_frames % 10 == 0:
                el = time.time() - t0
                print(f"[{name}] FPS: {10 / el:.2f}", flush=True)     # reference :320-328
                t0 = time.time()
                if count_:
                    print(f"[{name}] count: {dict(sorted(tracker.class_counts().items()))}", flush=True)
        if track_:
            rows = tracker.flush_pipelined()
            if rows is not None and pending:
                emit(labels, rows, tracker.last_det_index, pending.pop(0))
    if count_:
        print(f"[{name}] count: {dict(sorted(tracker.class_counts().items()))}")
    print(f"[{name}] {n_frames} frames, {n_rows} label rows -> {labels_path}")


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Process video with YOLO.")
    parser.add_argument("--source", nargs="+", type=str, default="0",
                        help="Input video file paths or camera indices")
    parser.add_argument("--track", action="store_true", help="if track objects")
    parser.add_argument("--count", action="store_true", help="if count objects")
    args = parser.parse_args()
    sources = args.source if isinstance(args.source, list) else [args.source]
    import torch
    n_gpu = max(torch.cuda.device_count(), 1)
    jobs = [{"source": s, "track": args.track, "count": args.count, "gpu": i % n_gpu}
            for i, s in enumerate(sources)]
    with get_context("spawn").Pool(processes=len(jobs)) as pool:     # reference :353-354
        pool.map(process_video, jobs)
