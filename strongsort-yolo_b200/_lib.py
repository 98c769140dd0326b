"""ctypes binding of libssb.so (include/ssb.h) and -- for the parity tests and tools only -- of
libssb_dbg.so (include/ssb_debug.h: the same code + the A/B baseline kernels and diagnostics).
No CPU fallback: a missing library or a missing CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSB_LIB") or os.path.join(HERE, "libssb.so")     # SSB_LIB: a tools/build_variants.py A/B build
LIB_DBG_PATH = os.path.join(HERE, "libssb_dbg.so")

# every symbol include/ssb.h declares (tests/test_abi.py checks the export list)
SYMBOLS = [
    "ssb_version", "ssb_launch_count", "ssb_last_error", "ssb_default_config", "ssb_workspace_bytes",
    "ssb_create", "ssb_destroy", "ssb_reset", "ssb_reid_tc_weight_bytes",
    "ssb_reid_set_weights_tc", "ssb_embed", "ssb_associate", "ssb_reid_tc_status", "ssb_update", "ssb_reid",
    "ssb_crop_boxes", "ssb_kf_predict", "ssb_kf_update", "ssb_kf_gating",
    "ssb_appearance_cost", "ssb_iou_cost", "ssb_lsap", "ssb_nms_scratch_bytes",
    "ssb_yolo_nms", "ssb_export_tracks",
    "ssb_yolo_num_anchors", "ssb_yolo_decode_v8", "ssb_yolo_decode_v5", "ssb_camera_update",
    "ssb_gallery_export", "ssb_gallery_cross_match", "ssb_gallery_cross_match_packed", "ssb_gallery_peer_match", "ssb_increment_ages",
    "ssb_profile_enable", "ssb_profile_read", "ssb_yolo_scale_boxes", "ssb_class_counts", "ssb_yolo_postprocess_v8",
    "ssb_appearance_tc_scratch_bytes", "ssb_appearance_cost_tc", "ssb_appearance_use_tc",
]

# every symbol include/ssb_debug.h adds (libssb_dbg.so only)
DEBUG_SYMBOLS = [
    "ssb_reid_num_tensors", "ssb_reid_tensor_sizes", "ssb_reid_set_weights", "ssb_reid_use_tc", "ssb_reid_block",
    "ssb_reid_tc_debug", "ssb_debug_cost_ptrs", "ssb_tc_probe",
]

SSB_CNT_N = 8
SSB_OUT_COLS = 8


class SsbConfig(C.Structure):
    _fields_ = [("max_tracks", C.c_int32), ("max_dets", C.c_int32),
                ("nn_budget", C.c_int32), ("feat_dim", C.c_int32),
                ("n_init", C.c_int32), ("max_age", C.c_int32),
                ("max_dist", C.c_double), ("max_iou_distance", C.c_double),
                ("mc_lambda", C.c_double), ("ema_alpha", C.c_double)]


class SsbError(RuntimeError):
    pass


_libs = {}


def load(debug=False):
    """Load libssb.so (or, with debug=True, libssb_dbg.so) once; raise loudly if it was not built."""
    if debug in _libs:
        return _libs[debug]
    path = LIB_DBG_PATH if debug else LIB_PATH
    if not os.path.exists(path):
        raise SsbError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.ssb_version.restype = i32
    lib.ssb_launch_count.restype = i64
    lib.ssb_last_error.restype = C.c_char_p
    lib.ssb_default_config.argtypes = [C.POINTER(SsbConfig)]
    lib.ssb_default_config.restype = None
    lib.ssb_workspace_bytes.argtypes = [C.POINTER(SsbConfig)]
    lib.ssb_workspace_bytes.restype = i64
    lib.ssb_create.argtypes = [C.POINTER(SsbConfig), vp, i64, C.POINTER(vp)]
    lib.ssb_destroy.argtypes = [vp]
    lib.ssb_reset.argtypes = [vp, vp]
    lib.ssb_increment_ages.argtypes = [vp, vp]
    lib.ssb_class_counts.argtypes = [vp, vp, vp]
    lib.ssb_reid_tc_weight_bytes.argtypes = [i32]
    lib.ssb_reid_tc_weight_bytes.restype = i64
    lib.ssb_reid_set_weights_tc.argtypes = [vp, vp, C.POINTER(i64), i32]
    lib.ssb_reid_tc_status.argtypes = [vp, C.POINTER(C.c_int32), vp]
    lib.ssb_embed.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, vp]
    lib.ssb_associate.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, i32, vp]
    lib.ssb_update.argtypes = [vp, vp, i32, vp, i32, i32, i32, vp, vp, vp, i32, vp]
    lib.ssb_reid.argtypes = [vp, vp, i32, i32, i32, vp, i32, vp, vp]
    lib.ssb_crop_boxes.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.ssb_kf_predict.argtypes = [vp, vp, i32, vp]
    lib.ssb_kf_update.argtypes = [vp, vp, vp, vp, i32, vp]
    lib.ssb_kf_gating.argtypes = [vp, vp, i32, vp, i32, vp, vp]
    lib.ssb_appearance_cost.argtypes = [vp, vp, i32, i32, vp, i32, i32, vp, vp]
    lib.ssb_profile_enable.argtypes = [vp, i32]
    lib.ssb_profile_read.argtypes = [vp, C.POINTER(C.c_float)]
    lib.ssb_appearance_tc_scratch_bytes.argtypes = [i32]
    lib.ssb_appearance_tc_scratch_bytes.restype = i64
    lib.ssb_appearance_cost_tc.argtypes = [vp, vp, i32, i32, vp, i32, i32, vp, vp, vp, vp]
    lib.ssb_appearance_use_tc.argtypes = [vp, i32]
    lib.ssb_iou_cost.argtypes = [vp, i32, vp, i32, vp, vp]
    lib.ssb_lsap.argtypes = [vp, i32, i32, vp, vp, vp]
    lib.ssb_nms_scratch_bytes.argtypes = [i32]
    lib.ssb_nms_scratch_bytes.restype = i64
    lib.ssb_yolo_nms.argtypes = [vp, i32, i32, i32, C.c_float, C.c_float, i32, i32, vp, vp, vp, vp]
    lib.ssb_yolo_scale_boxes.argtypes = [vp, i32, vp, i32, C.c_float, C.c_float, C.c_float, i32, i32, vp]
    lib.ssb_yolo_postprocess_v8.argtypes = [vp, i32, i32, i32, i32, C.c_float, C.c_float, i32, i32, C.c_float, C.c_float,
                                            C.c_float, i32, i32, vp, vp, vp, vp, vp]
    lib.ssb_yolo_num_anchors.argtypes = [i32, i32]
    lib.ssb_yolo_num_anchors.restype = i32
    lib.ssb_yolo_decode_v8.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    lib.ssb_yolo_decode_v5.argtypes = [vp, i32, i32, i32, C.c_float, vp, vp, vp]
    lib.ssb_camera_update.argtypes = [vp, C.POINTER(C.c_double), vp]
    lib.ssb_gallery_export.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.ssb_gallery_cross_match.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, C.c_float, vp, vp, vp, vp]
    lib.ssb_gallery_cross_match_packed.argtypes = [vp, i32, i32, i32, i32, C.c_float, vp, vp, vp, vp]
    lib.ssb_gallery_peer_match.argtypes = [vp, i32, i32, i32, i32, C.c_float, vp, vp, vp, vp, vp]
    lib.ssb_export_tracks.argtypes = [vp] * 11
    names = list(SYMBOLS)
    if debug:
        lib.ssb_reid_num_tensors.restype = i32
        lib.ssb_reid_tensor_sizes.argtypes = [C.POINTER(i64)]
        lib.ssb_reid_set_weights.argtypes = [vp, vp, C.POINTER(i64), i32]
        lib.ssb_reid_use_tc.argtypes = [vp, i32]
        lib.ssb_reid_block.argtypes = [vp, i32, vp, vp, i32, i32, vp]
        lib.ssb_reid_tc_debug.argtypes = [vp]
        lib.ssb_tc_probe.argtypes = [vp, i32, i32, vp, i32, i32, vp, vp, vp]
        lib.ssb_debug_cost_ptrs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        names += DEBUG_SYMBOLS
    for name in names:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("ssb_version", "ssb_reid_num_tensors", "ssb_yolo_num_anchors"):
            fn.restype = i32
    lib._ssb_debug = bool(debug)
    _libs[debug] = lib
    return lib


def last_error(lib=None):
    for l in ([lib] if lib is not None else list(_libs.values())):
        msg = l.ssb_last_error().decode("utf-8", "replace")
        if msg:
            return msg
    return ""


def check(rc, what=""):
    if rc != 0:
        raise SsbError(f"{what} failed (rc={rc}): {last_error()}")


def ptr(t):
    """device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise SsbError("strongsort_yolo_b200 needs a CUDA device (B200, sm_100a); "
                       "there is no CPU fallback")
    return torch
