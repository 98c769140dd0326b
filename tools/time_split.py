"""Time ssb_embed and ssb_associate separately (serial, CUDA events) on a workload:
  python tools/time_split.py C4 [frames]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strongsort_yolo_b200 import _lib, synth  # noqa: E402
from strongsort_yolo_b200.strong_sort import StrongSORT, _HDR_BYTES  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
    nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    lib = _lib.load(debug=True)
    trk = StrongSORT(max_tracks=2048, max_dets=640) if cfg == "C4" else StrongSORT(debug=True)
    st = synth.make_stream(cfg)
    frames = [st.next_frame() for _ in range(nfr)]
    imgs = [torch.from_numpy(f.img).cuda() for f in frames]
    dets = [torch.from_numpy(f.dets).cuda() for f in frames]
    H, W = frames[0].img.shape[:2]
    s = trk.stream
    sp = C.c_void_p(s.cuda_stream)
    out_rows = C.c_void_p(trk._out_dev.data_ptr() + _HDR_BYTES)
    hint = 0
    te, ta = [], []
    with torch.cuda.stream(s):
        for i in range(nfr):
            n = len(frames[i].dets)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(s)
            _lib.check(lib.ssb_embed(trk._h, 0, C.c_void_p(dets[i].data_ptr()), n, C.c_void_p(imgs[i].data_ptr()),
                                     H, W, 3 * W, sp))
            e[1].record(s)
            _lib.check(lib.ssb_associate(trk._h, 0, n, H, W, None, out_rows, C.c_void_p(trk._out_dev.data_ptr()),
                                         hint, sp))
            e[2].record(s)
            s.synchronize()
            cnt = trk._out_dev[:32].view(torch.int32).cpu().numpy()
            hint = int(cnt[1])
            te.append(e[0].elapsed_time(e[1])); ta.append(e[1].elapsed_time(e[2]))
            print(f"frame {i}: n={n} tracks={cnt[1]} confirmed={cnt[2]} embed {te[-1]:.3f} ms  associate {ta[-1]:.3f} ms", flush=True)
    k = nfr // 2
    print("median embed %.3f ms, associate %.3f ms" % (float(np.median(te[k:])), float(np.median(ta[k:]))))


if __name__ == "__main__":
    main()
