"""NumPy emulation of csrc/reid_tc.cu's index arithmetic (band split, padded
linear pixel maps, 9 shifted GEMMs, zero-ring masks, gate-scaled conv3, hi/lo
operands) driven by the SAME packed blob the kernel reads (weights.pack_tc).
It pins the host-side packing and every offset formula on the CPU; only the
PTX-level behaviour is left to the GPU tests."""
import numpy as np

from strongsort_yolo_b200 import weights


class Cfg:
    def __init__(self, b):
        cin, mid, midp, cout, down = weights.TC_BLOCKS[b]
        geo = [(64, 32, 16, 4, 4), (64, 32, 16, 4, 4), (32, 16, 8, 4, 4), (32, 16, 8, 4, 4),
               (16, 8, 16, 0, 1), (16, 8, 16, 0, 1)][b]
        self.CIN, self.MID, self.MIDP, self.COUT, self.DOWN = cin, mid, midp, cout, down
        self.H, self.W, self.R, self.HALO, self.NB = geo
        self.RH = self.R + 2 * self.HALO
        self.WP = self.W + 2
        self.NPX = (self.RH + 2) * self.WP
        self.NT = (self.NPX + 127) // 128
        self.GUARD = self.WP + 2
        self.MAP_PX = self.GUARD + self.NT * 128 + self.GUARD
        self.MCH = midp // 8
        self.IN_P0 = (1 + self.HALO) * self.WP
        self.IN_P1 = (1 + self.HALO + self.R) * self.WP
        self.IT0 = self.IN_P0 // 128
        self.IT1 = (self.IN_P1 + 127) // 128
        self.NIT = self.IT1 - self.IT0
        self.LCW_B = 2 * 9 * midp * midp * 2
        self.C1W_B = 2 * cin * midp * 2
        self.DNW_B = 2 * cin * cout * 2 if down else 0
        self.NPAR = midp + 10 * midp + cout + 2 * midp + 2 + 2 * midp + midp
        self.G_LCW = self.C1W_B + self.DNW_B
        self.G_PAR = self.G_LCW + 10 * self.LCW_B
        self.G_W3 = self.G_PAR + ((self.NPAR * 4 + 127) // 128) * 128


def _b_operand(raw, K, N):
    """bytes of [K/8][N][8] hi then lo -> float64 W[k][n] (hi + lo)."""
    half = K * N * 2
    hi = np.frombuffer(raw[:half], dtype=np.float16).reshape(K // 8, N, 8).astype(np.float64)
    lo = np.frombuffer(raw[half:2 * half], dtype=np.float16).reshape(K // 8, N, 8).astype(np.float64)
    return (hi + lo).transpose(0, 2, 1).reshape(K, N)


def _b_operand_cat(raw, K, N):
    """bytes of [K/8][N hi rows | N lo rows][8] -> float64 W[k][n]."""
    a = np.frombuffer(raw[:K * N * 4], dtype=np.float16).reshape(K // 8, 2 * N, 8).astype(np.float64)
    return (a[:, :N] + a[:, N:]).transpose(0, 2, 1).reshape(K, N)


def _hl(v):
    """operand rounding: fp32 -> fp16 hi + fp16 lo (as float64)."""
    v32 = np.asarray(v, dtype=np.float32)
    hi = v32.astype(np.float16)
    lo = (v32 - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64) + lo.astype(np.float64)


def osblock_emul(b, x, blob, offs):
    """x: float32 [n,H,W,cin] -> float32 [n,H,W,cout], band by band like the kernel."""
    c = Cfg(b)
    sec = bytes(blob[int(offs[b]):int(offs[b + 1])])
    bop = _b_operand_cat if weights.TC_CAT[b] else _b_operand
    W1 = bop(sec[0:c.C1W_B], c.CIN, c.MIDP)
    WD = _b_operand(sec[c.C1W_B:c.C1W_B + c.DNW_B], c.CIN, c.COUT) if c.DOWN else None
    LCW = [bop(sec[c.G_LCW + l * c.LCW_B:c.G_LCW + (l + 1) * c.LCW_B], 9 * c.MIDP, c.MIDP)
           for l in range(10)]
    par = np.frombuffer(sec[c.G_PAR:c.G_PAR + c.NPAR * 4], dtype=np.float32).astype(np.float64)
    o = 0
    b1 = par[o:o + c.MIDP]; o += c.MIDP
    blc = par[o:o + 10 * c.MIDP].reshape(10, c.MIDP); o += 10 * c.MIDP
    b3 = par[o:o + c.COUT]; o += c.COUT
    gw1 = par[o:o + 2 * c.MIDP].reshape(c.MIDP, 2); o += 2 * c.MIDP
    gb1 = par[o:o + 2]; o += 2
    gw2 = par[o:o + 2 * c.MIDP].reshape(2, c.MIDP); o += 2 * c.MIDP
    gb2 = par[o:o + c.MIDP]; o += c.MIDP
    W3 = np.frombuffer(sec[c.G_W3:c.G_W3 + c.MIDP * c.COUT * 4], dtype=np.float32).reshape(c.MIDP, c.COUT).astype(np.float64)

    n = x.shape[0]
    y = np.zeros((n, c.H, c.W, c.COUT), dtype=np.float32)
    P = np.arange(c.NT * 128)
    LR, LC_ = P // c.WP, P % c.WP
    for crop in range(n):
        bands = []
        for band in range(c.NB):
            row0 = band * c.R - c.HALO
            GR, GC = row0 + LR - 1, LC_ - 1
            valid = (P < c.NPX) & (LC_ >= 1) & (LC_ <= c.W) & (LR >= 1) & (LR <= c.RH) & (GR >= 0) & (GR < c.H)
            own = valid & (LR >= 1 + c.HALO) & (LR < 1 + c.HALO + c.R)
            xt = np.zeros((c.NT * 128, c.CIN))
            xt[valid] = x[crop, GR[valid], GC[valid]]
            xt = _hl(xt)                                            # staged operand tiles
            acc1 = xt @ W1                                          # conv1, every band tile
            accd = (xt @ WD) if c.DOWN else None                    # downsample (inner tiles used)

            def to_map(acc, bias):
                f = np.where(valid[:, None], np.maximum(acc + bias, 0.0), 0.0).astype(np.float32)
                m = np.full((c.MAP_PX, c.MIDP), np.nan)             # guards: garbage
                m[c.GUARD:c.GUARD + c.NT * 128] = _hl(f)
                return m, f

            X1, _ = to_map(acc1, b1)
            bands.append(dict(valid=valid, own=own, GR=GR, GC=GC, X1=X1, accd=accd, streams=[]))
        # streams: all bands of the crop advance together (cluster).  P/Q are persistent
        # per-band maps; a LightConv with `rem` successors only updates the tiles
        # [tile_lo(rem), tile_hi(rem)) ("trapezoid"), the rest keeps stale content.
        def t_lo(r):
            return (max(0, 1 + c.HALO - r) * c.WP) // 128

        def t_hi(r):
            row = min(1 + c.HALO + c.R + r, c.RH + 2)
            return min(c.NT, (row * c.WP + 127) // 128)

        for bd in bands:
            bd["P"] = np.full((c.MAP_PX, c.MIDP), 7.25)      # stale / uninitialised content
            bd["Q"] = np.full((c.MAP_PX, c.MIDP), -3.5)
        c3 = [np.zeros((c.NT * 128, c.COUT)) + (bd["accd"] if c.DOWN else 0.0) for bd in bands]
        lc = 0
        for s in range(4):
            src_name, dst_name = "X1", "P"
            fs = None
            for k in range(s + 1):
                rem = s - k
                lo, hi = t_lo(rem) * 128, t_hi(rem) * 128
                fs = []
                for bd in bands:
                    src = bd[src_name]
                    acc = np.zeros((c.NT * 128, c.MIDP))
                    for tap in range(9):
                        off = (tap // 3 - 1) * c.WP + (tap % 3 - 1)
                        a = src[c.GUARD + off:c.GUARD + off + c.NT * 128]
                        a = np.where(np.isnan(a), 1e30, a)          # guard garbage must never matter
                        acc += a @ LCW[lc][tap * c.MIDP:(tap + 1) * c.MIDP]
                    valid = bd["valid"]
                    f = np.where(valid[:, None], np.maximum(acc + blc[lc], 0.0), 0.0).astype(np.float32)
                    bd[dst_name][c.GUARD + lo:c.GUARD + hi] = _hl(f)[lo:hi]
                    fs.append(f)
                src_name, dst_name = dst_name, ("Q" if dst_name == "P" else "P")
                lc += 1
            srcs = [bd[src_name] for bd in bands]
            tot = sum(f[bd["own"]].astype(np.float64).sum(0) for f, bd in zip(fs, bands))
            mean = tot / (c.H * c.W)
            h = np.maximum(gb1 + mean @ gw1, 0.0)
            g = 1.0 / (1.0 + np.exp(-(gb2 + h @ gw2)))
            W3g = _hl(W3 * g[:, None])
            for i, (bd, m) in enumerate(zip(bands, srcs)):
                a = m[c.GUARD:c.GUARD + c.NT * 128].copy()
                a[:c.IT0 * 128] = 0.0                                # conv3 only runs on the inner tiles
                a[c.IT1 * 128:] = 0.0
                c3[i] += a @ W3g
        for bd, acc in zip(bands, c3):
            own = bd["own"].copy()
            own[:c.IT0 * 128] = False
            own[c.IT1 * 128:] = False
            assert own.sum() == bd["own"].sum(), "inner tiles do not cover the band's own pixels"
            out = acc[own] + b3
            if not c.DOWN:
                out = out + x[crop, bd["GR"][own], bd["GC"][own]]
            y[crop, bd["GR"][own], bd["GC"][own]] = np.maximum(out, 0.0)
    return y


# ---------------------------------------------------------------------------
# csrc/reid_tc3.cu: pointwise on the tensor cores (hi/lo operands), depthwise in fp32
# ---------------------------------------------------------------------------
class Cfg3:
    GEO = [(64, 32, 16, 4, 4, 4), (64, 32, 16, 4, 4, 4), (32, 16, 8, 4, 4, 5), (32, 16, 8, 4, 4, 5),
           (16, 8, 16, 0, 1, 8), (16, 8, 16, 0, 1, 8)]

    def __init__(self, b):
        cin, mid, midp, cout, down = weights.TC_BLOCKS[b]
        self.CIN, self.MID, self.MIDP, self.COUT, self.DOWN = cin, mid, midp, cout, down
        self.H, self.W, self.R, self.HALO, self.NB, self.SEG = self.GEO[b]
        self.RH = self.R + 2 * self.HALO
        self.NPX = self.RH * self.W
        assert self.NPX % 128 == 0
        self.NT = self.NPX // 128
        self.OWN_P0, self.OWN_P1 = self.HALO * self.W, (self.HALO + self.R) * self.W
        self.IT0, self.IT1 = self.OWN_P0 // 128, (self.OWN_P1 + 127) // 128
        self.C1W_B = cin * midp * 4
        self.DNW_B = cin * cout * 4 if down else 0
        self.LCW_B = midp * midp * 4
        self.WALL_B = self.C1W_B + self.DNW_B + 10 * self.LCW_B
        self.NPAR = midp + 100 * midp + cout + 2 * midp + 2 + 2 * midp + midp
        self.G_PAR = (self.WALL_B + 127) // 128 * 128
        self.G_W3 = self.G_PAR + (self.NPAR * 4 + 127) // 128 * 128
        self.G_TOTAL = self.G_W3 + midp * cout * 4


def osblock3_emul(b, x, blob, offs):
    """x: float32 [n,H,W,cin] -> float32 [n,H,W,cout]; band by band, in-place P map with stale
    rows outside each layer's trapezoid, exactly the kernel's index arithmetic."""
    c = Cfg3(b)
    end = int(offs[11 + b]) if 11 + b < len(offs) else len(blob)
    sec = bytes(blob[int(offs[10 + b]):end])
    assert len(sec) >= c.G_TOTAL
    W1 = _b_operand_cat(sec[0:c.C1W_B], c.CIN, c.MIDP)
    WD = _b_operand(sec[c.C1W_B:c.C1W_B + c.DNW_B], c.CIN, c.COUT) if c.DOWN else None
    o0 = c.C1W_B + c.DNW_B
    PW = [_b_operand_cat(sec[o0 + l * c.LCW_B:o0 + (l + 1) * c.LCW_B], c.MIDP, c.MIDP) for l in range(10)]
    par = np.frombuffer(sec[c.G_PAR:c.G_PAR + c.NPAR * 4], dtype=np.float32)
    o = 0
    b1 = par[o:o + c.MIDP].astype(np.float64); o += c.MIDP
    DW, BL = [], []
    for _ in range(10):
        DW.append(par[o:o + 9 * c.MIDP].reshape(9, c.MIDP)); o += 9 * c.MIDP
        BL.append(par[o:o + c.MIDP]); o += c.MIDP
    b3 = par[o:o + c.COUT].astype(np.float64); o += c.COUT
    gw1 = par[o:o + 2 * c.MIDP].reshape(c.MIDP, 2).astype(np.float64); o += 2 * c.MIDP
    gb1 = par[o:o + 2].astype(np.float64); o += 2
    gw2 = par[o:o + 2 * c.MIDP].reshape(2, c.MIDP).astype(np.float64); o += 2 * c.MIDP
    gb2 = par[o:o + c.MIDP].astype(np.float64); o += c.MIDP
    assert o == c.NPAR
    W3 = np.frombuffer(sec[c.G_W3:c.G_W3 + c.MIDP * c.COUT * 4], dtype=np.float32).reshape(c.MIDP, c.COUT).astype(np.float64)

    n = x.shape[0]
    y = np.zeros((n, c.H, c.W, c.COUT), dtype=np.float32)
    Pix = np.arange(c.NPX)
    LR, COL = Pix // c.W, Pix % c.W
    for crop in range(n):
        bands = []
        for band in range(c.NB):
            row0 = band * c.R - c.HALO
            GR = row0 + LR
            valid = (GR >= 0) & (GR < c.H)
            xt = np.zeros((c.NPX, c.CIN))
            xt[valid] = x[crop, GR[valid], COL[valid]]
            xt = _hl(xt)
            X1 = _hl(np.where(valid[:, None], np.maximum(xt @ W1 + b1, 0.0), 0.0).astype(np.float32))
            accd = (xt @ WD) if c.DOWN else np.zeros((c.NPX, c.COUT))
            Pm = np.zeros((c.NPX, c.MIDP))                        # zeroed after phase 1
            bands.append(dict(row0=row0, valid=valid, GR=GR, X1=X1, P=Pm, c3=accd.copy()))
        lc = 0
        for s in range(4):
            for k in range(s + 1):
                rem = s - k
                ra, rb = max(c.HALO - rem, 0), min(c.HALO + c.R + rem, c.RH)
                fs = []
                for bd in bands:
                    src = bd["X1"] if k == 0 else bd["P"]
                    T = np.zeros((c.RH + 2, c.W + 2, c.MIDP), dtype=np.float32)       # zero ring
                    T[1:-1, 1:-1] = (src @ PW[lc]).astype(np.float32).reshape(c.RH, c.W, c.MIDP)
                    out = np.zeros((c.RH, c.W, c.MIDP), dtype=np.float32)
                    out[:] = BL[lc]
                    for dy in range(3):
                        for dx in range(3):
                            out += DW[lc][dy * 3 + dx] * T[dy:dy + c.RH, dx:dx + c.W]
                    rows_in = ((bd["row0"] + np.arange(c.RH)) >= 0) & ((bd["row0"] + np.arange(c.RH)) < c.H)
                    out = np.where(rows_in[:, None, None], np.maximum(out, 0.0), 0.0).astype(np.float32)
                    newP = bd["P"].reshape(c.RH, c.W, c.MIDP).copy()
                    newP[ra:rb, :, :c.MID] = _hl(out[ra:rb, :, :c.MID])    # only the trapezoid rows, real channels
                    bd["P"] = newP.reshape(c.NPX, c.MIDP)
                    fs.append(out)
                lc += 1
            tot = sum(f[c.HALO:c.HALO + c.R, :, :].astype(np.float64).sum((0, 1)) for f in fs)
            tot[c.MID:] = 0.0
            mean = tot / (c.H * c.W)
            h = np.maximum(gb1 + mean @ gw1, 0.0)
            g = 1.0 / (1.0 + np.exp(-(gb2 + h @ gw2)))
            W3g = _hl((W3 * g[:, None]).astype(np.float32))
            for bd in bands:
                a = bd["P"].copy()
                a[:c.IT0 * 128] = 0.0
                a[c.IT1 * 128:] = 0.0
                bd["c3"] += a @ W3g
        for band, bd in enumerate(bands):
            own = (Pix >= c.OWN_P0) & (Pix < c.OWN_P1)
            assert own[c.IT0 * 128:c.IT1 * 128].sum() == own.sum()
            out = bd["c3"][own] + b3
            gr, gc = bd["GR"][own], COL[own]
            if not c.DOWN:
                out = out + x[crop, gr, gc]
            y[crop, gr, gc] = np.maximum(out, 0.0)
    return y
