"""OSNet-x0.25 checkpoint loading, BatchNorm folding and packing for the device.

Input: a torchreid-style state_dict ({name: array}; .npz, or a torch .pth/.pt
checkpoint).  Output: one flat float32 blob whose tensor order and layouts are
the canonical walk of csrc/reid.cu (``reid_tensor_sizes_host``):

  stem  W[ky][kx][ci][co], b      (conv1 7x7 + BN folded)
  per OSBlock: conv1 W[ci][co], b | 10 x LightConv (pw W[ci][co], dw W[tap][c]
               with BN scale folded, b) in order a1,b1,b2,c1..c3,d1..d4 |
               gate fc1 W[c][r], b, fc2 W[r][c], b | conv3 W[ci][co], b |
               downsample W[ci][co], b (if cin != cout) | transition W, b
  conv5 W[ci][co], b ; fc W[ci][co] (BN1d folded), b
Each tensor starts on a 16-byte boundary inside the blob.
"""
from __future__ import annotations

import os

import numpy as np

BN_EPS = 1e-5
BLOCKS = [(16, 64), (64, 64), (64, 96), (96, 96), (96, 128), (128, 128)]
DEFAULT_WEIGHTS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "weights",
                               "osnet_x0_25_synth.npz")


def load_state_dict(path=None):
    path = path or DEFAULT_WEIGHTS
    if path.endswith(".npz"):
        return {k: np.asarray(v) for k, v in np.load(path).items()}
    import torch
    sd = torch.load(path, map_location="cpu")
    if isinstance(sd, dict) and "state_dict" in sd:
        sd = sd["state_dict"]
    return {k.replace("module.", "", 1): v.detach().cpu().numpy() for k, v in sd.items()
            if hasattr(v, "detach")}


def _bn(sd, prefix):
    g = sd[prefix + ".weight"].astype(np.float64)
    b = sd[prefix + ".bias"].astype(np.float64)
    m = sd[prefix + ".running_mean"].astype(np.float64)
    v = sd[prefix + ".running_var"].astype(np.float64)
    scale = g / np.sqrt(v + BN_EPS)
    return scale, b - m * scale


def _fold_conv(sd, conv, bn):
    """[co,ci,kh,kw] conv + BN -> W[kh][kw][ci][co] (squeezed for 1x1), b[co]."""
    w = sd[conv + ".weight"].astype(np.float64)
    scale, shift = _bn(sd, bn)
    w = w * scale[:, None, None, None]
    w = np.transpose(w, (2, 3, 1, 0))                     # kh, kw, ci, co
    if w.shape[0] == 1 and w.shape[1] == 1:
        w = w[0, 0]                                       # [ci][co]
    return w, shift


def fold(sd):
    """-> list of (name, float32 ndarray) in canonical order."""
    out = []
    w, b = _fold_conv(sd, "conv1.conv", "conv1.bn")
    out += [("stem.w", w), ("stem.b", b)]
    stage_of = ["conv2.0", "conv2.1", "conv3.0", "conv3.1", "conv4.0", "conv4.1"]
    streams = [["conv2a"], ["conv2b.0", "conv2b.1"],
               ["conv2c.0", "conv2c.1", "conv2c.2"],
               ["conv2d.0", "conv2d.1", "conv2d.2", "conv2d.3"]]
    for bi, (cin, cout) in enumerate(BLOCKS):
        p = stage_of[bi]
        w, b = _fold_conv(sd, f"{p}.conv1.conv", f"{p}.conv1.bn")
        out += [(f"{p}.conv1.w", w), (f"{p}.conv1.b", b)]
        for names in streams:
            for nm in names:
                q = f"{p}.{nm}"
                pw = sd[q + ".conv1.weight"].astype(np.float64)[:, :, 0, 0].T   # [ci][co]
                scale, shift = _bn(sd, q + ".bn")
                dw = sd[q + ".conv2.weight"].astype(np.float64)[:, 0]          # [c][3][3]
                dw = (dw * scale[:, None, None]).reshape(dw.shape[0], 9).T     # [tap][c]
                out += [(q + ".pw", pw), (q + ".dw", dw), (q + ".b", shift)]
        g = f"{p}.gate"
        out += [(g + ".fc1.w", sd[g + ".fc1.weight"].astype(np.float64)[:, :, 0, 0].T),
                (g + ".fc1.b", sd[g + ".fc1.bias"].astype(np.float64)),
                (g + ".fc2.w", sd[g + ".fc2.weight"].astype(np.float64)[:, :, 0, 0].T),
                (g + ".fc2.b", sd[g + ".fc2.bias"].astype(np.float64))]
        w, b = _fold_conv(sd, f"{p}.conv3.conv", f"{p}.conv3.bn")
        out += [(f"{p}.conv3.w", w), (f"{p}.conv3.b", b)]
        if cin != cout:
            w, b = _fold_conv(sd, f"{p}.downsample.conv", f"{p}.downsample.bn")
            out += [(f"{p}.down.w", w), (f"{p}.down.b", b)]
        if bi in (1, 3):
            t = p.split(".")[0] + ".2.0"
            w, b = _fold_conv(sd, f"{t}.conv", f"{t}.bn")
            out += [(f"{t}.w", w), (f"{t}.b", b)]
    w, b = _fold_conv(sd, "conv5.conv", "conv5.bn")
    out += [("conv5.w", w), ("conv5.b", b)]
    scale, shift = _bn(sd, "fc.1")
    fw = sd["fc.0.weight"].astype(np.float64) * scale[:, None]            # [co][ci]
    fb = sd["fc.0.bias"].astype(np.float64) * scale + shift
    out += [("fc.w", fw.T), ("fc.b", fb)]
    return [(n, np.ascontiguousarray(a, dtype=np.float32)) for n, a in out]


def pack(tensors):
    """-> (blob float32 [total], sizes int64 [n]) with 16-byte aligned tensors."""
    sizes = np.asarray([a.size for _, a in tensors], dtype=np.int64)
    padded = (sizes + 3) & ~3
    blob = np.zeros(int(padded.sum()), dtype=np.float32)
    off = 0
    for (_, a), p in zip(tensors, padded):
        blob[off:off + a.size] = a.reshape(-1)
        off += int(p)
    return blob, sizes


# ---------------------------------------------------------------------------
# tensor-core operand blob (csrc/reid_tc.cu: BlkCfg / Par)
# ---------------------------------------------------------------------------
#            CIN MID MIDP COUT DOWN
TC_BLOCKS = [(16, 16, 16, 64, True), (64, 16, 16, 64, False), (64, 24, 32, 96, True),
             (96, 24, 32, 96, False), (96, 32, 32, 128, True), (128, 32, 32, 128, False)]
# BlkCfg::CAT of csrc/reid_tc.cu (TMEM has room for 2*MIDP columns per 3x3 tile): the conv1
# and LightConv weights are then laid out [k/8][hi rows | lo rows][8] instead of hi block, lo block
TC_CAT = [False, False, True, True, True, True]
_STAGE = ["conv2.0", "conv2.1", "conv3.0", "conv3.1", "conv4.0", "conv4.1"]
_LC = ["conv2a", "conv2b.0", "conv2b.1", "conv2c.0", "conv2c.1", "conv2c.2",
       "conv2d.0", "conv2d.1", "conv2d.2", "conv2d.3"]


def _hi_lo(a64):
    """fp32 value -> (hi, lo) fp16 pair, hi + lo == fp32(a) to ~2^-22."""
    a32 = np.asarray(a64, dtype=np.float32)
    hi = a32.astype(np.float16)
    lo = (a32 - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def _b_layout(w_kn):
    """B operand, K-major no-swizzle: W[k][n] -> bytes of [k/8][n][8] hi then lo."""
    K, N = w_kn.shape
    assert K % 8 == 0
    t = np.ascontiguousarray(w_kn.reshape(K // 8, 8, N).transpose(0, 2, 1))   # [kc][n][8]
    hi, lo = _hi_lo(t)
    return hi.tobytes() + lo.tobytes()


def _b_layout_cat(w_kn):
    """N-concatenated B operand: W[k][n] -> bytes of [k/8][n (hi) then n (lo)][8]."""
    K, N = w_kn.shape
    t = np.ascontiguousarray(w_kn.reshape(K // 8, 8, N).transpose(0, 2, 1))   # [kc][n][8]
    hi, lo = _hi_lo(t)
    return np.concatenate([hi, lo], axis=1).tobytes()


def _pad128(b):
    return b + b"\0" * ((-len(b)) % 128)


def _pack_tc3_block(T, bi):
    """One OSBlock section of csrc/reid_tc3.cu (B3::G_*): pointwise weights as N-concatenated
    hi/lo tensor-core operands, depthwise taps + biases in fp32."""
    cin, mid, midp, cout, down = TC_BLOCKS[bi]
    p = _STAGE[bi]
    w = np.zeros((cin, midp)); w[:, :mid] = T[f"{p}.conv1.w"]
    sec = _b_layout_cat(w)                                                     # C1W
    if down:
        sec += _b_layout(T[f"{p}.down.w"].astype(np.float64))                  # DNW [cin][cout]
    for nm in _LC:                                                             # LCW x 10 (pointwise)
        pw = np.zeros((midp, midp)); pw[:mid, :mid] = T[f"{p}.{nm}.pw"]
        sec += _b_layout_cat(pw)
    sec = _pad128(sec)
    par = np.zeros(midp + 100 * midp + cout + 2 * midp + 2 + 2 * midp + midp, dtype=np.float32)
    o = 0
    par[o:o + mid] = T[f"{p}.conv1.b"]; o += midp
    for nm in _LC:
        dw = np.zeros((9, midp), dtype=np.float32); dw[:, :mid] = T[f"{p}.{nm}.dw"]     # [tap][c]
        par[o:o + 9 * midp] = dw.reshape(-1); o += 9 * midp
        par[o:o + mid] = T[f"{p}.{nm}.b"]; o += midp
    b3 = T[f"{p}.conv3.b"].astype(np.float64)
    if down:
        b3 = b3 + T[f"{p}.down.b"].astype(np.float64)
    par[o:o + cout] = b3; o += cout
    g1 = T[f"{p}.gate.fc1.w"]                                                  # [mid][r]
    r = g1.shape[1]
    gw1 = np.zeros((midp, 2), dtype=np.float32); gw1[:mid, :r] = g1
    par[o:o + 2 * midp] = gw1.reshape(-1); o += 2 * midp
    par[o:o + r] = T[f"{p}.gate.fc1.b"]; o += 2
    gw2 = np.zeros((2, midp), dtype=np.float32); gw2[:r, :mid] = T[f"{p}.gate.fc2.w"]
    par[o:o + 2 * midp] = gw2.reshape(-1); o += 2 * midp
    par[o:o + mid] = T[f"{p}.gate.fc2.b"]; o += midp
    assert o == par.size
    sec += _pad128(par.tobytes())
    w3 = np.zeros((midp, cout), dtype=np.float32); w3[:mid] = T[f"{p}.conv3.w"]
    return sec + w3.tobytes()


def pack_tc(tensors):
    """tensors: output of fold().  -> (uint8 blob, int64 offsets[16]): sections 0..5 OSBlocks
    for reid_tc.cu (9 shifted GEMMs per LightConv), 6..7 transitions, 8 tail, 9 stem,
    10..15 OSBlocks for reid_tc3.cu (pointwise on tcgen05 + fp32 depthwise)."""
    T = dict(tensors)
    blob, offs = b"", []
    for bi, (cin, mid, midp, cout, down) in enumerate(TC_BLOCKS):
        p = _STAGE[bi]
        sec = b""
        lay = _b_layout_cat if TC_CAT[bi] else _b_layout
        w = np.zeros((cin, midp)); w[:, :mid] = T[f"{p}.conv1.w"]
        sec += lay(w)                                                          # C1W
        if down:
            sec += _b_layout(T[f"{p}.down.w"].astype(np.float64))              # DNW [cin][cout]
        for nm in _LC:                                                         # LCW x 10
            q = f"{p}.{nm}"
            pw = T[q + ".pw"].astype(np.float64)                               # [ci][c]
            dw = T[q + ".dw"].astype(np.float64)                               # [tap][c]
            wd = np.zeros((9, midp, midp))                                     # [tap][ci][c]
            wd[:, :mid, :mid] = dw[:, None, :] * pw[None, :, :]
            sec += lay(wd.reshape(9 * midp, midp))
        par = np.zeros(midp + 10 * midp + cout + 2 * midp + 2 + 2 * midp + midp, dtype=np.float32)
        o = 0
        par[o:o + mid] = T[f"{p}.conv1.b"]; o += midp
        for nm in _LC:
            par[o:o + mid] = T[f"{p}.{nm}.b"]; o += midp
        b3 = T[f"{p}.conv3.b"].astype(np.float64)
        if down:
            b3 = b3 + T[f"{p}.down.b"].astype(np.float64)
        par[o:o + cout] = b3; o += cout
        g1 = T[f"{p}.gate.fc1.w"]                                              # [mid][r]
        r = g1.shape[1]
        gw1 = np.zeros((midp, 2), dtype=np.float32); gw1[:mid, :r] = g1
        par[o:o + 2 * midp] = gw1.reshape(-1); o += 2 * midp
        par[o:o + r] = T[f"{p}.gate.fc1.b"]; o += 2
        gw2 = np.zeros((2, midp), dtype=np.float32); gw2[:r, :mid] = T[f"{p}.gate.fc2.w"]
        par[o:o + 2 * midp] = gw2.reshape(-1); o += 2 * midp
        par[o:o + mid] = T[f"{p}.gate.fc2.b"]; o += midp
        assert o == par.size
        sec = sec + _pad128(par.tobytes())
        w3 = np.zeros((midp, cout), dtype=np.float32); w3[:mid] = T[f"{p}.conv3.w"]
        sec += w3.tobytes()
        offs.append(len(blob))
        blob += _pad128(sec)
    # transition layers (conv1x1 + ReLU + avgpool): W hi/lo | bias
    for t in ("conv2.2.0", "conv3.2.0"):
        offs.append(len(blob))
        blob += _pad128(_b_layout(T[t + ".w"].astype(np.float64)) + _pad128(T[t + ".b"].astype(np.float32).tobytes()))
    # tail: conv5 W hi/lo | b5 | fc W [128][512] fp32 | fc b
    offs.append(len(blob))
    blob += _pad128(_b_layout(T["conv5.w"].astype(np.float64)) + T["conv5.b"].astype(np.float32).tobytes()
                    + np.ascontiguousarray(T["fc.w"], dtype=np.float32).tobytes()
                    + T["fc.b"].astype(np.float32).tobytes())
    # stem as a 4x4 conv on the 2x2 space-to-depth image: per tap (a,b) a [16 k][16 co] matrix,
    # k = (dy*2+dx)*3 + c  <-  W[2a+dy][2b+dx][c][co]   (zero where 2a+dy or 2b+dx > 6, k >= 12)
    sw = T["stem.w"].astype(np.float64)                       # [7][7][3][16]
    wt = np.zeros((16, 16, 16))
    for a in range(4):
        for b in range(4):
            for dy in range(2):
                for dx in range(2):
                    if 2 * a + dy <= 6 and 2 * b + dx <= 6:
                        e = (dy * 2 + dx) * 3
                        wt[a * 4 + b, e:e + 3, :] = sw[2 * a + dy, 2 * b + dx]
    hi, lo = _hi_lo(np.ascontiguousarray(wt.reshape(16, 2, 8, 16).transpose(0, 1, 3, 2)))   # [tap][kc][n][8]
    offs.append(len(blob))
    cat = np.concatenate([hi, lo], axis=2)                    # [tap][kc][hi 16 rows | lo 16 rows][8]
    blob += _pad128(cat.tobytes() + _pad128(T["stem.b"].astype(np.float32).tobytes()))
    for bi in range(6):
        offs.append(len(blob))
        blob += _pad128(_pack_tc3_block(T, bi))
    return np.frombuffer(blob, dtype=np.uint8).copy(), np.asarray(offs, dtype=np.int64)
