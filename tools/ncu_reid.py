"""One ReID forward (all kernels of ssb_reid) on a C2 frame, for `ncu --set full`:

  ncu --set full --clock-control none --import-source on --profile-from-start off \
      -o gpurun_out/prof_reid python tools/ncu_reid.py

2 warm forwards, then one forward between cudaProfilerStart/Stop (the only one ncu captures).
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strongsort_yolo_b200 import _lib, synth  # noqa: E402
from strongsort_yolo_b200.strong_sort import StrongSORT  # noqa: E402


def main():
    backend = sys.argv[1] if len(sys.argv) > 1 else "tc"
    dbg = backend != "tc"                       # the product library for the product path, libssb_dbg.so for the baselines
    lib = _lib.load(debug=dbg)
    P = lambda t: C.c_void_p(t.data_ptr())
    trk = StrongSORT(debug=dbg)
    st = synth.make_stream("C2")
    fr = [st.next_frame() for _ in range(3)][-1]
    n = len(fr.dets)
    img = torch.from_numpy(fr.img).cuda()
    dets = torch.from_numpy(fr.dets).cuda()
    boxes = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    feats = torch.zeros((n, 512), dtype=torch.float32, device="cuda")
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.ssb_crop_boxes(P(dets), n, 1080, 1920, P(boxes), sp))
    trk.set_reid_backend(backend)
    for r in range(3):
        if r == 2:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        _lib.check(lib.ssb_reid(trk._h, P(img), 1080, 1920, 1920 * 3, P(boxes), n, P(feats), sp))
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("crops", n, "status", trk.reid_tc_status())


if __name__ == "__main__":
    main()
