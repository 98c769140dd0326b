#!/usr/bin/env python
"""bench.py -- tracked frames/sec of the StrongSORT per-frame path on B200.

Metric (BASELINE.json): tracked frames/sec at 1080p, 100 dets/frame (config C2:
"1 stream 1080p synthetic, 100 dets/frame, OSNet-x0.25 ReID on 1xB200"), one
independent synthetic video stream per GPU (stream i -> GPU i, no data-path
collective: weak scaling).

One step = one frame through ``StrongSORT.update(dets, img)``: OSNet ReID of
every detection crop + Kalman predict + gated appearance cost + IoU cost + two
linear assignments + track-table update.  Seeded random-weight OSNet (no
pretrained file exists offline) and synthetic frames -- "data": "synthetic".

  value  : frames/s with frames and detections already resident in HBM, through the
           two-stage pipeline (``update_pipelined``: the OSNet of frame k on one stream
           overlaps the association of frame k-1 on another; identical results), K frames
           timed as a whole with CUDA events; inputs rotate through W+K distinct frames
           (>= 5x the L2), ``config.serial_flushed_ms_per_step`` is the one-frame-at-a-
           time figure with a 256 MiB L2 flush between steps
  e2e    : frames/s through the public synchronous ``StrongSORT.update`` with HOST inputs
           (pinned frame + dets), H2D and the D2H of the result rows inside
           the timed region, one synchronisation per frame
  roofline     : ReID forward (``ssb_reid``, the dominant kernels) timed alone with CUDA
                 events; algorithmic flops 2*82.3e6*N per frame vs the measured
                 bf16 peak in MEASURED_PEAKS.json; traffic = DRAM bytes of the same
                 forward from the committed ncu --set full capture
  cpu_baseline : the CPU oracle (oracle/: NumPy/SciPy tracker + fp32 torch
                 OSNet; the reference's own StrongSORT code is absent) on the
                 host cores over a bounded sample of the same stream
  clocks       : SM clock / throttle reasons polled through NVML during the timed region

``--workload C4`` runs BASELINE.json's configs[3]; ``--shared-gallery`` adds config C5's
per-frame cross-stream exchange (NCCL all-gather) to the timed region.
``--impl reference`` times that CPU oracle alone and prints the same line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "tracked frames/sec at 1080p, 100 dets/frame"
UNIT = "frames/s"
REID_MACS_PER_CROP = 82314880          # oracle/osnet_torch.count_macs()
WORKLOAD = "C2: 1 stream/GPU, 1080p synthetic, 100 dets/frame, OSNet-x0.25 ReID + StrongSORT"
# --workload C4 (not the default bench line): BASELINE.json configs[3]
WORKLOADS = {
    "C2": (METRIC, WORKLOAD),
    "C4": ("tracked frames/sec at 4K, 500 dets/frame, 256 live tracks",
           "C4: 1 stream/GPU, 4K synthetic crowded scene, 500 dets/frame, 256 persistent tracks, "
           "OSNet-x0.25 ReID + StrongSORT"),
}
CFG = "C2"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def gen_frames(stream_id, count, pin):
    """Pre-generate `count` frames of the C2 stream (host arrays; pinned torch
    tensors when `pin`)."""
    import torch
    from strongsort_yolo_b200 import synth
    st = synth.make_stream(CFG, stream_id=stream_id)
    imgs, dets = [], []
    for _ in range(count):
        fr = st.next_frame()
        t = torch.from_numpy(fr.img)
        imgs.append(t.pin_memory() if pin else t)
        dets.append(fr.dets)
    return imgs, dets


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: an NVML polling thread (every 20 ms,
    in-process, so even a 150 ms region gets several samples); falls back to `nvidia-smi -lms` if NVML
    cannot be loaded.  CUDA_VISIBLE_DEVICES is honoured through the device's UUID / PCI bus id."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = False
        self._thr = None
        self._smi = None
        self._nv = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            props = torch.cuda.get_device_properties(gpu_index)
            h = None
            try:                                   # the CUDA device's UUID survives CUDA_VISIBLE_DEVICES remapping
                u = str(props.uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID((u if u.startswith("GPU-") else "GPU-" + u).encode())
            except Exception:
                h = None
            if h is None:
                try:
                    h = pynvml.nvmlDeviceGetHandleByPciBusId(
                        ("%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)).encode())
                except Exception:
                    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                    ids = [int(v) for v in vis.split(",") if v.strip().isdigit()]
                    h = pynvml.nvmlDeviceGetHandleByIndex(ids[gpu_index] if gpu_index < len(ids) else gpu_index)
            self._h = h
            self._nv = pynvml
        except Exception:
            self._nv = None

    def _poll(self):
        nv = self._nv
        bits = ((nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap"))
        while not self._stop:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                for bit, name in bits:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self._nv is not None:
            import threading
            try:
                self.max_mhz = float(self._nv.nvmlDeviceGetMaxClockInfo(self._h, self._nv.NVML_CLOCK_SM))
            except Exception:
                self.max_mhz = None
            self._thr = threading.Thread(target=self._poll, daemon=True)
            self._thr.start()
            return
        try:
            self._f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self._smi = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=self._f, stderr=subprocess.DEVNULL)
        except Exception:
            self._smi = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": None}
        if self._thr is not None:
            self._stop = True
            self._thr.join(timeout=2)
            src = "nvml"
            if not self.samples:                   # the poller never ran inside the region: one reading right after it
                try:
                    self.samples.append(float(self._nv.nvmlDeviceGetClockInfo(self._h, self._nv.NVML_CLOCK_SM)))
                    src = "nvml (single reading right after the timed region)"
                except Exception:
                    pass
            if self.samples:
                out.update(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.max_mhz,
                           reasons=sorted(self.reasons), samples=len(self.samples), source=src)
            return out
        if self._smi is None:
            return out
        self._smi.terminate()
        try:
            self._smi.wait(timeout=5)
        except Exception:
            self._smi.kill()
        self._f.flush()
        rows = [r.strip().split(",") for r in open(self._f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self._f.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); out["sm_max_mhz"] = float(r[2])
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        out["source"] = "nvidia-smi"
        return out


def pick_threads(img, dets):
    """More threads is not faster for this oracle (small convs: 128 threads ran
    60x slower than 16 on the GPU box's host).  Time one ReID batch at a few
    thread counts and keep the fastest -- "all the host threads it can use"."""
    import torch
    from oracle import osnet_torch, strongsort_np
    from strongsort_yolo_b200 import weights
    ext = osnet_torch.OracleExtractor(weights.load_state_dict(), batch=128)
    xywh = strongsort_np.xyxy2xywh(dets[:, :4])
    boxes = np.asarray([strongsort_np.crop_box_xyxy(b, img.shape[1], img.shape[0]) for b in xywh])[:32]
    n = os.cpu_count() or 1
    best, best_t = 1, float("inf")
    for th in sorted({c for c in (4, 8, 16, 32, 64, n) if c <= n}):
        torch.set_num_threads(th)
        ext(img, boxes[:4])
        t0 = time.perf_counter()
        ext(img, boxes)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = th, dt
        if dt > 4 * best_t:
            break
    return best


def run_oracle(frames_img, frames_dets, warm, timed, threads):
    """CPU oracle fps over frames [warm, warm+timed) after `warm` untimed frames."""
    import torch
    from oracle import osnet_torch, strongsort_np
    from strongsort_yolo_b200 import weights
    torch.set_num_threads(threads)
    ora = strongsort_np.StrongSORTOracle(
        osnet_torch.OracleExtractor(weights.load_state_dict(), batch=128))
    per = []
    for i in range(warm + timed):
        img = frames_img[i].numpy() if hasattr(frames_img[i], "numpy") else frames_img[i]
        t0 = time.perf_counter()
        ora.update(frames_dets[i], img)
        dt = time.perf_counter() - t0
        if i >= warm:
            per.append(dt)
    return per


def impl_reference(args, rank):
    """The reference arm: the CPU oracle on the host cores (the reference's own
    StrongSORT code is absent from /root/reference -- SURVEY.md section 0)."""
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    steps, warm = min(args.steps, 60), min(args.warmup, 5)
    from strongsort_yolo_b200 import synth
    st = synth.make_stream(CFG, stream_id=0)
    imgs, dets = [], []
    for _ in range(warm + steps):
        fr = st.next_frame()
        imgs.append(fr.img); dets.append(fr.dets)
    cores = pick_threads(imgs[0], dets[0])
    per = run_oracle(imgs, dets, warm, steps, cores)
    ms = 1000.0 * float(np.mean(per))
    fps = 1000.0 / ms
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU restatement of StrongSORT (reference code absent); "
                   "1 stream on the host cores regardless of --gpus"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": int(torch.get_num_threads()),
                         "kind": "port", "sample": f"{steps} frames after {warm} warm-up frames of the C2 stream"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=20)
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS),
                    help="C2 (default, the BASELINE.json metric) or C4 (4K, 500 dets/frame)")
    ap.add_argument("--shared-gallery", action="store_true",
                    help="config C5's optional exchange: all-gather every stream's confirmed-track features over "
                         "NCCL after each frame and match across streams (read-only), inside the timed region")
    ap.add_argument("--only-device", action="store_true",
                    help="profiling aid: run only the device-resident timed loop (for ncu launch lists)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    global CFG, METRIC, WORKLOAD
    CFG = args.workload
    METRIC, WORKLOAD = WORKLOADS[CFG]
    if args.impl == "reference":
        impl_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    from strongsort_yolo_b200 import dist as ssb_dist
    if world > 1:
        ssb_dist.init("nccl", device)
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ROOT, "strongsort-yolo_b200", "libssb.so")):
        if local_rank == 0:
            ge.build()
        if world > 1:
            dist.barrier()
    import ctypes as C
    from strongsort_yolo_b200 import _lib
    from strongsort_yolo_b200.strong_sort import StrongSORT, _HDR_BYTES
    lib = _lib.load()
    peaks, peak_src = load_peaks()

    K, W = args.steps, max(args.warmup, 3)
    total = W + K
    imgs, dets = gen_frames(rank, total, pin=True)
    n_per = [len(d) for d in dets]
    H, Wd = imgs[0].shape[0], imgs[0].shape[1]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        return ssb_dist.max_over_ranks(x, device)

    # ---------------- value: device-resident inputs, per-step events ----------
    trk_kw = dict(max_tracks=2048, max_dets=640) if CFG == "C4" else {}
    trk = StrongSORT(device=str(device), **trk_kw)
    st = trk.stream
    imgs_dev = [im.to(device) for im in imgs]
    dets_dev = [torch.from_numpy(d).to(device) for d in dets]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    out_rows = C.c_void_p(trk._out_dev.data_ptr() + _HDR_BYTES)
    sptr = C.c_void_p(st.cuda_stream)
    hint = 0

    def step_device(i):
        nonlocal hint
        _lib.check(lib.ssb_update(trk._h, _lib.ptr(dets_dev[i]), n_per[i], _lib.ptr(imgs_dev[i]), H, Wd,
                                  3 * Wd, None, out_rows, _lib.ptr(trk._out_dev), hint, sptr), "ssb_update")

    def read_hint():
        nonlocal hint
        st.synchronize()
        hint = int(trk._out_dev[:32].view(torch.int32)[1].item())

    # (a) serial reference: one frame at a time, per-step events, L2 flushed between steps
    with torch.cuda.stream(st):
        for i in range(W):
            step_device(i)
            read_hint()
    barrier()
    KS = min(K, 30)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(KS)]
    with torch.cuda.stream(st):
        for k in range(KS):
            flush.fill_(k & 0xFF)                      # L2 flush, outside the timed pair
            ev[k][0].record(st)
            step_device(W + k)
            ev[k][1].record(st)
            read_hint()                                # sizes the next frame's grids (exact)
    barrier()
    serial_ms = sum(a.elapsed_time(b) for a, b in ev) / KS
    del trk

    # (b) value: the two-stage pipeline (embedding of frame k on one stream overlaps the
    # association of frame k-1 on another), device-resident inputs, K frames timed as a whole.
    # Inputs rotate through W+K distinct frames (>= 5x the 126 MB L2), no flush needed.
    trk = StrongSORT(device=str(device), **trk_kw)
    st = trk.stream
    sptr = C.c_void_p(st.cuda_stream)
    gal = ssb_dist.SharedGallery(trk) if args.shared_gallery else None
    for i in range(W):
        trk.update_pipelined(dets_dev[i], imgs_dev[i])
        if gal is not None:
            gal.step()
    trk.flush_pipelined()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.ssb_launch_count()
    e0.record(st)
    for k in range(K):
        trk.update_pipelined(dets_dev[W + k], imgs_dev[W + k])
        if gal is not None:
            gal.step()                     # export + NCCL all-gather + cross-stream match, on the tracker's stream
    trk.flush_pipelined()
    if gal is not None:
        st.wait_stream(gal.stream)         # the exchange of the last frame ends inside the timed region
    e1.record(st)
    barrier()
    launches = int(lib.ssb_launch_count() - l0)
    t_steps_ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    t_steps_ms = max_over_ranks(t_steps_ms)
    ms_per_step = t_steps_ms / K
    value = world * K / (t_steps_ms / 1000.0)
    final_next_id = int(trk.last_counts[3])

    if args.only_device:
        if rank == 0:
            print(json.dumps({"only_device": True, "ms_per_step": ms_per_step, "value": value,
                              "serial_flushed_ms_per_step": serial_ms,
                              "gpu_launches": launches, "steps": K, "warmup": W}), flush=True)
        return

    # ---------------- ReID alone: roofline of the dominant kernels ------------
    i0 = W + K // 2
    boxes = torch.zeros((n_per[i0], 4), dtype=torch.int32, device=device)
    feats = torch.zeros((n_per[i0], 512), dtype=torch.float32, device=device)
    with torch.cuda.stream(st):
        _lib.check(lib.ssb_crop_boxes(_lib.ptr(dets_dev[i0]), n_per[i0], H, Wd, _lib.ptr(boxes), sptr))
        reps = 20
        rev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for r in range(3 + reps):
            if r >= 3:
                flush.fill_(r)
                rev[r - 3][0].record(st)
            _lib.check(lib.ssb_reid(trk._h, _lib.ptr(imgs_dev[i0]), H, Wd, 3 * Wd, _lib.ptr(boxes),
                                    n_per[i0], _lib.ptr(feats), sptr))
            if r >= 3:
                rev[r - 3][1].record(st)
    st.synchronize()
    reid_ms = float(np.mean([a.elapsed_time(b) for a, b in rev]))
    reid_flops = 2.0 * REID_MACS_PER_CROP * n_per[i0]
    achieved_tf = reid_flops / (reid_ms * 1e-3) / 1e12
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "reid_traffic_bytes.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_reid_forward")
        except Exception:
            traffic = None

    # ---------------- e2e: host buffers through StrongSORT.update -------------
    n_cross = len(gal.report()) if gal is not None else None
    trk2 = StrongSORT(device=str(device), **trk_kw)
    for i in range(W):
        trk2.update(dets[i], imgs[i])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(trk2.stream)
    t0 = time.perf_counter()
    for k in range(K):
        rows = trk2.update(dets[W + k], imgs[W + k])
    e1.record(trk2.stream)
    barrier()
    e2e_s = max(time.perf_counter() - t0, e0.elapsed_time(e1) / 1000.0)
    e2e_s = max_over_ranks(e2e_s)
    e2e_fps = world * K / e2e_s
    same_ids = int(trk2.last_counts[3]) == final_next_id
    h2d = int(imgs[0].numel() + np.mean(n_per) * 24)
    d2h = int(trk2._out_bytes)

    # ---------------- CPU baseline (rank 0, N=1 only) --------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        img0 = imgs[0].numpy()
        cores = pick_threads(img0, dets[0])
        warm_c, timed_c = 5, args.cpu_frames
        per = run_oracle(imgs, dets, warm_c, timed_c, cores)
        cpu = {"value": 1.0 / float(np.mean(per)), "unit": UNIT, "cores": int(torch.get_num_threads()),
               "kind": "port", "host_cores": os.cpu_count(),
               "sample": f"{timed_c} frames after {warm_c} warm-up frames of the same C2 stream, "
                         "thread count = fastest of {4,8,16,32,64,all}; "
                         "CPU restatement of StrongSORT (reference code absent from the snapshot)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (ReID, appearance) + f64 (Kalman, gating, LSAP)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "streams": world, "frame": [H, Wd, 3],
                       "dets_per_frame": float(np.mean(n_per)),
                       "l2": f"not flushed: inputs rotate through {W + K} distinct frames "
                             f"({(W + K) * imgs[0].numel() / 1e6:.0f} MB > 126 MB L2)",
                       "pipeline": "embedding(k) on stream 1 overlaps association(k-1) on stream 2 "
                                   "(ssb_embed / ssb_associate); results identical to the serial path",
                       "serial_flushed_ms_per_step": serial_ms,
                       "weights": "seeded random OSNet-x0.25, BN calibrated on synthetic crops",
                       "e2e_ids_equal_device_run": bool(same_ids),
                       **({"shared_gallery": "per frame: export + all-gather (NCCL) of [256,512] f32 + ids per rank + "
                                             "cross-stream cosine match, read-only, on a side stream; inside the timed region",
                           "cross_stream_matches_last_frame_rank0": n_cross} if gal is not None else {})},
            "e2e": {"value": e2e_fps, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": "OSNet ReID forward (all kernels of ssb_reid)",
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": traffic, "peak_source": peak_src,
                         "reid_ms": reid_ms, "flops_per_launch": reid_flops},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
