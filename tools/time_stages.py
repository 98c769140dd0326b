"""Time the stage entry points of the C-ABI alone with CUDA events (B200):
ReID forward (tc / simt), appearance cost, LSAP on realistic clamped matrices,
gating, IoU.  Prints a JSON dict of microseconds per call (warm, L2 not flushed).

  python tools/time_stages.py > gpurun_out/stages.json
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strongsort_yolo_b200 import _lib, synth  # noqa: E402
from strongsort_yolo_b200.strong_sort import StrongSORT  # noqa: E402


def timeit(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
    return {"median_us": t[len(t) // 2], "min_us": t[0]}


def main():
    lib = _lib.load(debug=True)           # baselines + phase stamps: libssb_dbg.so
    P = lambda t: C.c_void_p(t.data_ptr())
    ST = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {}
    trk = StrongSORT(debug=True)
    st = synth.make_stream("C2")
    frames = [st.next_frame() for _ in range(8)]
    for fr in frames:
        trk.update(fr.dets, fr.img)
    fr = frames[-1]
    n = len(fr.dets)
    img = torch.from_numpy(fr.img).cuda()
    dets = torch.from_numpy(fr.dets).cuda()
    boxes = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    feats = torch.zeros((n, 512), dtype=torch.float32, device="cuda")
    _lib.check(lib.ssb_crop_boxes(P(dets), n, 1080, 1920, P(boxes), ST()))
    for backend in ("tc", "tc3", "tc9", "simt"):
        trk.set_reid_backend(backend)
        out[f"reid_{backend}_n{n}"] = timeit(lambda: _lib.check(lib.ssb_reid(
            trk._h, P(img), 1080, 1920, 1920 * 3, P(boxes), n, P(feats), ST())))
    trk.set_reid_backend("tc")
    # phase stamps of CTA 0 inside every tensor-core OSBlock (cycles since kernel start)
    dbg = torch.zeros(128, dtype=torch.int64, device="cuda")
    shapes = [(64, 32, 16), (64, 32, 64), (32, 16, 64), (32, 16, 96), (16, 8, 96), (16, 8, 128)]
    for b, (hh, ww, cin) in enumerate(shapes):
        x = np.maximum(np.random.default_rng(b).normal(0.5, 1, (n, hh, ww, cin)), 0).astype(np.float32)
        for mode, tag in ((3, ""), (2, "_tc3"), (1, "_tap9")):
            trk.reid_block(b, x, mode)
            lib.ssb_reid_tc_debug(P(dbg))
            trk.reid_block(b, x, mode)
            lib.ssb_reid_tc_debug(None)
            torch.cuda.synchronize()
            d = dbg.cpu().numpy()
            k = int(d[0])
            out[f"osblock{b}{tag}_phase_cycles"] = [int(v - d[1]) for v in d[2:1 + k]]
            if mode == 3 and int(d[64]) > 0:     # fine stamps of layers 1 and 2 (reid_tc4.cu fstamp), deltas
                f = [int(v) for v in d[65:65 + int(d[64])]]
                out[f"osblock{b}_fine_deltas"] = [f[i + 1] - f[i] for i in range(len(f) - 1)]
            dbg.zero_()
    # stem phases (CTA 0): S built, MMAs done, conv drained, pooled
    lib.ssb_reid_tc_debug(P(dbg))
    _lib.check(lib.ssb_reid(trk._h, P(img), 1080, 1920, 1920 * 3, P(boxes), n, P(feats), ST()))
    torch.cuda.synchronize()
    lib.ssb_reid_tc_debug(None)
    d = dbg.cpu().numpy()
    out["stem_phase_cycles"] = [int(v - d[41]) for v in d[42:41 + int(d[40])]]
    # stage-A-like clamped cost matrices from the live tracker
    a, b = trk.debug_costs()
    # phase stamps of the last assign_stage_a_kernel (debug build): cycles
    pa, pb, pd = C.c_void_p(), C.c_void_p(), C.c_void_p()
    _lib.check(lib.ssb_debug_cost_ptrs(trk._h, C.byref(pa), C.byref(pb), C.byref(pd)))
    from strongsort_yolo_b200.strong_sort import _wrap_device
    cnt = _wrap_device(torch, pd.value, (20,), "<i4", trk.device).cpu().numpy()     # cnt[FC_ROWS_A ...]
    out["assign_stage_a_cycles"] = {"lsap_block": int(cnt[10]), "lists": int(cnt[11]), "staged_at": int(cnt[12]),
                                    "solved_at": int(cnt[13]), "search_steps": int(cnt[14]),
                                    "search_cycles": int(cnt[15]), "other_cycles": int(cnt[16])}
    for name, m in (("lsap_stageA", a), ("lsap_stageB", b)):
        if m.size:
            md = torch.as_tensor(m).cuda()
            c4r = torch.zeros(m.shape[0], dtype=torch.int32, device="cuda")
            r4c = torch.zeros(m.shape[1], dtype=torch.int32, device="cuda")
            out[f"{name}_{m.shape[0]}x{m.shape[1]}"] = timeit(lambda: _lib.check(lib.ssb_lsap(
                P(md), m.shape[0], m.shape[1], P(c4r), P(r4c), ST())))
    rng = np.random.default_rng(0)
    for (nr, nc) in ((100, 100), (256, 500)):
        m = rng.random((nr, nc)) * 0.3
        m[m > 0.2] = 0.2 + 1e-5
        md = torch.as_tensor(m).cuda()
        c4r = torch.zeros(nr, dtype=torch.int32, device="cuda")
        r4c = torch.zeros(nc, dtype=torch.int32, device="cuda")
        out[f"lsap_random_clamped_{nr}x{nc}"] = timeit(lambda: _lib.check(lib.ssb_lsap(
            P(md), nr, nc, P(c4r), P(r4c), ST())), reps=10, warm=2)
    for (T, N) in ((100, 100), (256, 500)):
        gal = torch.rand((T, 100, 512), device="cuda")
        cnt = torch.full((T,), 100, dtype=torch.int32, device="cuda")
        f = torch.rand((N, 512), device="cuda")
        o = torch.zeros((T, N), device="cuda")
        out[f"appearance_T{T}_N{N}"] = timeit(lambda: _lib.check(lib.ssb_appearance_cost(
            P(gal), P(cnt), T, 100, P(f), N, 512, P(o), ST())))
    # whole update with device-resident inputs (no flush)
    hint = [int(trk.last_counts[1])]
    outp = C.c_void_p(trk._out_dev.data_ptr() + 64)

    def upd():
        _lib.check(lib.ssb_update(trk._h, P(dets), n, P(img), 1080, 1920, 1920 * 3, None, outp,
                                  P(trk._out_dev), -1, ST()))
    out["ssb_update_same_frame"] = timeit(upd, reps=20, warm=3)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
