"""Build A/B variants of libssb.so (compile-time band geometry of the OSBlock kernels, csrc/reid_tc4.cu) into
strongsort-yolo_b200/variants/ and, on the GPU box, time the ReID forward of each:

  python tools/build_variants.py build                   # here (nvcc cross-compiles)
  python tools/build_variants.py time > gpurun_out/variants.json     # on the B200 box
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# name: (extra nvcc flags, environment of the timing run)
VARIANTS = {
    "base": ([], {}),                                       # the default build
    "base_nopdl": ([], {"SSB_PDL": "0"}),                   # programmatic dependent launch off
    "base_nosplit": ([], {"SSB_SPLIT": "0"}),               # whole frame on one stream instead of two halves
    "base_split2": ([], {"SSB_SPLIT": "2"}),                # two / four parts on two / four streams (default: three)
    "base_split4": ([], {"SSB_SPLIT": "4"}),
    "base_nopwf": ([], {"SSB_PW_FUSED": "0"}),              # transitions as separate pw_tc launches (11 instead of 9)
    "dw1chain": (["-DSSB_DW_CHAINS=1"], {}),                # single 9-term depthwise chain (4 fewer instructions per row)
    "s3split": (["-DSSB_S3_SPLIT=true"], {}),
    "s2r16": (["-DSSB_S2_R=16", "-DSSB_S2_SPLIT=true"], {}),
    "aptbig": (["-DSSB_APT_BIG"], {}),                      # appearance_tc with 160-192 KB stages (round-2 first version)
    "stem8": (["-DSSB_STEM_PR=8"], {}),                     # stem with 8 pooled rows per CTA, 512 threads, one CTA per SM
}


def build():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_b", os.path.join(ROOT, "strongsort-yolo_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    built = set()
    for name, (flags, _env) in VARIANTS.items():
        key = name.split("_")[0]
        if key not in built:
            print(b.build_variant(key, flags))
            built.add(key)


def time_one():
    """(child process, SSB_LIB set) ReID forward of one C2 frame: median us over 30 warm launches + parity vs the KAT."""
    import ctypes as C
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    from strongsort_yolo_b200 import _lib, synth
    from strongsort_yolo_b200.strong_sort import StrongSORT
    lib = _lib.load()
    trk = StrongSORT()
    g = np.load(os.path.join(ROOT, "tests", "golden", "reid_kat.npz"))
    emb = trk.extract_features(g["img"], g["boxes"])
    ref = g["emb"]
    err = float(np.max(np.abs(emb - ref) / np.abs(ref).max(axis=1, keepdims=True)))
    fr = synth.make_stream("C2").next_frame()
    n = len(fr.dets)
    img = torch.from_numpy(fr.img).cuda()
    dets = torch.from_numpy(fr.dets).cuda()
    boxes = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    feats = torch.zeros((n, 512), dtype=torch.float32, device="cuda")
    P = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.ssb_crop_boxes(P(dets), n, 1080, 1920, P(boxes), st))
    fn = lambda: _lib.check(lib.ssb_reid(trk._h, P(img), 1080, 1920, 1920 * 3, P(boxes), n, P(feats), st))
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
    print(json.dumps({"median_us": t[len(t) // 2], "min_us": t[0], "kat_err": err, "status": trk.reid_tc_status(), "n": n}))


def time_all():
    out = {}
    for name, (_flags, extra) in VARIANTS.items():
        lib = os.path.join(ROOT, "strongsort-yolo_b200", "variants", f"libssb_{name.split('_')[0]}.so")
        if not os.path.exists(lib):
            continue
        env = dict(os.environ, SSB_LIB=lib, **extra)
        try:
            p = subprocess.run([sys.executable, __file__, "time_one"], env=env, capture_output=True, text=True, timeout=240)
            out[name] = json.loads(p.stdout.strip().splitlines()[-1]) if p.returncode == 0 else {"error": p.stderr[-400:]}
        except Exception as e:          # a hung variant must not take the others down
            out[name] = {"error": repr(e)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"build": build, "time": time_all, "time_one": time_one}[sys.argv[1] if len(sys.argv) > 1 else "build"]()
