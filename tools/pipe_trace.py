"""GPU timeline of the two-stage pipeline (StrongSORT.update_pipelined, lag 2): when the embedding and the association
of consecutive frames start and end on the device, from CUDA events recorded around ssb_embed / ssb_associate.
    python tools/pipe_trace.py [C2|C4] [frames]
Prints one row per frame (ms since the first traced frame's embedding started) + the mean durations."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    bench.CFG = cfg
    from strongsort_yolo_b200.strong_sort import StrongSORT
    dev = torch.device("cuda", 0)
    W = 6
    imgs, dets, _raws = bench.gen_frames(0, W + frames, pin=False)
    imgs_dev = [im.to(dev) for im in imgs]
    dets_dev = [torch.from_numpy(d.astype(np.float32)).to(dev) for d in dets]
    kw = dict(max_tracks=2048, max_dets=640) if cfg == "C4" else {}
    trk = StrongSORT(device=str(dev), **kw)
    for i in range(W):
        trk.update_pipelined(dets_dev[i], imgs_dev[i], lag=2)
    trk.pipeline_trace = []
    for i in range(W, W + frames):
        trk.update_pipelined(dets_dev[i], imgs_dev[i], lag=2)
    trk.drain_pipelined()
    torch.cuda.synchronize()
    tr = trk.pipeline_trace
    base = tr[0][1][0]
    rows = []
    for k, ev in tr:
        t = [base.elapsed_time(e) for e in ev]
        rows.append({"frame": k, "embed_start": t[0], "embed_end": t[1], "assoc_start": t[2], "assoc_end": t[3]})
    emb = np.array([r["embed_end"] - r["embed_start"] for r in rows])
    asc = np.array([r["assoc_end"] - r["assoc_start"] for r in rows])
    period = (rows[-1]["assoc_end"] - rows[4]["assoc_end"]) / (len(rows) - 5)
    out = {"config": cfg, "frames": frames, "embed_ms_mean": float(emb[4:].mean()), "assoc_ms_mean": float(asc[4:].mean()),
           "period_ms": float(period), "rows": rows}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
