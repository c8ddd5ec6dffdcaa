"""Synthetic post-parse content: there is no VVC bitstream or encoder on the box, so the
work lists a flattener would emit from VVdeC's parsed Picture are drawn from a seeded generator instead.
Everything here is host-side input preparation (numpy); no pixel arithmetic of the hot path lives here."""
import numpy as np
from . import abi


def partition(rng, W, H, ctu=128, min_dim=4, min_area=32, p_split=None):
    """Random QT+BT partition of a WxH picture into CUs. Returns int array [n,4] = x,y,w,h (luma).
    Blocks crossing the picture boundary are split until they fit (implicit boundary splits)."""
    out = []
    p_split = p_split or {128: 0.94, 64: 0.68, 32: 0.6, 16: 0.6, 8: 0.5, 4: 0.0}

    def rec(x, y, w, h):
        if x >= W or y >= H:
            return
        cross_x, cross_y = x + w > W, y + h > H
        big = max(w, h)
        if cross_x or cross_y:
            if cross_x and cross_y or (w == h and w > 64):
                mode = "q"
            else:
                mode = "v" if cross_x else "h"
        else:
            r = rng.random()
            if r >= p_split.get(big, 0.0):
                out.append((x, y, w, h)); return
            opts = []
            if w == h and w >= 16: opts += ["q", "q"]
            if w >= 2 * min_dim and (w // 2) * h >= min_area and w <= 64: opts.append("v")
            if h >= 2 * min_dim and w * (h // 2) >= min_area and h <= 64: opts.append("h")
            if not opts:
                out.append((x, y, w, h)); return
            mode = opts[rng.integers(0, len(opts))]
        if mode == "q":
            hw, hh = w // 2, h // 2
            for dy in (0, hh):
                for dx in (0, hw):
                    rec(x + dx, y + dy, hw, hh)
        elif mode == "v":
            rec(x, y, w // 2, h); rec(x + w // 2, y, w // 2, h)
        else:
            rec(x, y, w, h // 2); rec(x, y + h // 2, w, h // 2)

    for cy in range(0, H, ctu):
        for cx in range(0, W, ctu):
            rec(cx, cy, ctu, ctu)
    return np.array(out, np.int32).reshape(-1, 4)


def _laplace_levels(rng, n, heavy=False):
    v = rng.laplace(0, 2.0 if not heavy else 400.0, size=n)
    return np.clip(np.rint(v), -32768, 32767).astype(np.int16)


def gen_tus(rng, cus, bit_depth=10, p_cbf=0.5, p_mts=0.2, p_lfnst=0.1, p_ts=0.05, p_bdpcm=0.03, p_jccr=0.1,
            p_full=0.3, chroma=True, heavy=0.02, dep_quant=True, p_intra=0.15, scaling=None):
    """TU records + packed level arena for a CU list (one TU per <=64x64 tile of each CU, all 3 components).
    Mirrors what the flattener (vvdec_glue/flatten_tu.h) derives from parsed TUs; QP drawn uniformly in 22..37.
    scaling: None, or the dict of gen_scaling_lists(): non-TS, non-LFNST TUs then use explicit scaling lists (Quant.cpp:309-312):
    B200_TU_SCALING, slOff -> the table of their (log2w, log2h), right shift + 4."""
    recs, coefs = [], []
    ncoef = 0
    inv_scales = np.array([[40, 45, 51, 57, 64, 72], [57, 64, 72, 80, 90, 102]])
    for (cx, cy, cw, ch) in cus:
        if rng.random() >= p_cbf:
            continue
        qp = int(rng.integers(22, 38)) + 6 * (bit_depth - 8)
        intra = rng.random() < p_intra
        for ty in range(cy, cy + ch, 64):
            for tx in range(cx, cx + cw, 64):
                tw, th = min(64, cw), min(64, ch)
                jccr = chroma and rng.random() < p_jccr
                for comp in ((0, 1, 2) if chroma else (0,)):
                    w, h = (tw, th) if comp == 0 else (tw >> 1, th >> 1)
                    x, y = (tx, ty) if comp == 0 else (tx >> 1, ty >> 1)
                    if min(w, h) < 2 or (comp and rng.random() < 0.3 and not jccr):
                        continue
                    ict = 0
                    if comp and jccr:
                        if comp == 2: continue
                        ict = int(rng.choice([-3, -2, -1, 1, 2, 3]))
                        if abs(ict) == 3: comp = 2
                    l2w, l2h = int(np.log2(w)), int(np.log2(h))
                    flags, tr, lfnst = 0, 0, 0
                    r = rng.random()
                    maxX = maxY = None
                    if w <= 32 and h <= 32 and r < p_bdpcm and intra:
                        flags = abi.TU_TS | (abi.TU_BDPCM_H if rng.random() < 0.5 else abi.TU_BDPCM_V)
                        maxX, maxY = w - 1, h - 1
                    elif w <= 32 and h <= 32 and r < p_bdpcm + p_ts:
                        flags = abi.TU_TS
                    elif comp == 0 and min(w, h) >= 4 and max(w, h) <= 32 and r < p_bdpcm + p_ts + p_mts:
                        th_, tv_ = rng.choice([abi.TR_DST7, abi.TR_DCT8], size=2)
                        tr = int(th_) | (int(tv_) << 2)
                    elif intra and min(w, h) >= 4 and r < p_bdpcm + p_ts + p_mts + p_lfnst:
                        lf_set = int(rng.integers(0, 4))
                        lfnst = int(rng.integers(1, 3)) | (lf_set << 2) | ((int(rng.integers(0, 2)) if lf_set else 0) << 4)
                    if maxX is None:
                        limx = min(w, 32) if not (tr & 3 and w == 32) else 16
                        limy = min(h, 32) if not ((tr >> 2) & 3 and h == 32) else 16
                        if lfnst:
                            # at most 16 (8 for 4x4/8x8) levels in scan order, all inside the first 4x4 coefficient group
                            maxX, maxY = 3, 3
                        elif rng.random() < p_full:
                            maxX, maxY = limx - 1, limy - 1
                        else:
                            maxX = min(limx - 1, int(rng.geometric(0.25)) - 1)
                            maxY = min(limy - 1, int(rng.geometric(0.25)) - 1)
                        if not (flags & abi.TU_TS) and not (maxX == 0 and maxY == 0):
                            # the parser reports the coded corner in whole coefficient groups (CABACReader.cpp:2447-2452): 4x4, or
                            # 2x8 / 8x2 for 2-wide / 2-high blocks; only a lone DC coefficient gives (0,0)
                            cgw, cgh = (2, 8) if w == 2 else (8, 2) if h == 2 else (4, 4)
                            maxX = min(limx, (maxX // cgw + 1) * cgw) - 1
                            maxY = min(limy, (maxY // cgh + 1) * cgh) - 1
                    is_ts = bool(flags & abi.TU_TS)
                    sqrt2 = (not is_ts) and ((l2w + l2h) & 1)
                    dq = dep_quant and not is_ts
                    q = max(qp, 4) if is_ts else qp
                    per, rem = ((q + 1) // 6, (q + 1) % 6) if dq else (q // 6, q % 6)
                    tr_shift = 15 - bit_depth - ((l2w + l2h) >> 1) + (-1 if sqrt2 else 0)
                    right_shift = 6 + (1 if dq else 0) - ((0 if is_ts else tr_shift) + per)
                    sl_off = 0
                    if scaling is not None and not is_ts and not lfnst:
                        flags |= abi.TU_SCALING; right_shift += 4; sl_off = scaling["off"][(l2w, l2h)]
                    in_bits = min(16, 32 + right_shift - 7)
                    n = (maxX + 1) * (maxY + 1)
                    lv = _laplace_levels(rng, n, rng.random() < heavy)
                    if lfnst:
                        # zero everything outside the first 8 scan positions (x+y<=2 covers 6 of them: safe subset)
                        yy, xx = np.divmod(np.arange(n), maxX + 1)
                        lv[(xx + yy) > 2] = 0
                    if lv[-1] == 0: lv[-1] = 1
                    recs.append((x, y, l2w, l2h, comp, flags, maxX, maxY, tr, lfnst, ict, right_shift, in_bits,
                                 int(inv_scales[1 if sqrt2 else 0][rem]), ncoef, sl_off, (0, 0)))
                    coefs.append(lv); ncoef += n
    tus = np.array(recs, dtype=abi.TU_DTYPE) if recs else np.zeros(0, abi.TU_DTYPE)
    arena = np.concatenate(coefs) if coefs else np.zeros(0, np.int16)
    return tus, arena


def gen_scaling_lists(rng):
    """One dequantisation table per block shape, laid out as Quant::getDequantCoeff returns them (w*h int32, value 16 = neutral), back to
    back in one arena; 'off' maps (log2w, log2h) to the table's offset.  Low frequencies get smaller values, as typical lists do."""
    off, parts, o = {}, [], 0
    for l2w in range(1, 7):
        for l2h in range(1, 7):
            w, h = 1 << l2w, 1 << l2h
            yy, xx = np.mgrid[0:h, 0:w]
            t = 8 + ((xx * 8) // w + (yy * 8) // h) * int(rng.integers(1, 12)) + rng.integers(0, 4, size=(h, w))
            off[(l2w, l2h)] = o; parts.append(np.clip(t, 1, 255).astype(np.int32).reshape(-1)); o += w * h
    return dict(off=off, arena=np.concatenate(parts))


def noise_planes(rng, W, H, bit_depth=10, chroma=True, strides=None):
    """Stand-in prediction / reference pictures: smooth gradient + band-limited noise, clipped to the bit depth."""
    mx = (1 << bit_depth) - 1
    out = []
    for c in range(3 if chroma else 1):
        w, h = (W, H) if c == 0 else (W >> 1, H >> 1)
        st = strides[c] if strides else w
        yy, xx = np.mgrid[0:h, 0:w]
        base = (mx / 2) + (mx / 4) * np.sin(xx / (37.0 + 11 * c)) * np.cos(yy / (53.0 - 7 * c))
        n = rng.normal(0, mx / 40, size=(h // 4 + 2, w // 4 + 2))
        n = np.kron(n, np.ones((4, 4)))[:h, :w]
        fine = rng.integers(-8, 9, size=(h, w))
        p = np.zeros((h, st), np.int16)
        p[:, :w] = np.clip(base + n + fine, 0, mx).astype(np.int16)
        out.append(p)
    return out


LF_DTYPE = np.dtype([("qp", "i1", (3,)), ("bs", "u1"), ("len", "u1"), ("flags", "u1")])
LFSLICE_DTYPE = np.dtype([("beta", "i1", (3,)), ("tc", "i1", (3,)), ("disable", "u1"), ("rsv", "u1")])


def unit_maps(cus, W, H, cu_attr=None):
    """Paint per-4x4-unit maps from a CU list: TU id, TU width/height (TUs are the <=64 tiles of a CU), CU index."""
    W4, H4 = (W + 3) // 4, (H + 3) // 4
    tu_id = np.zeros((H4, W4), np.int32); tu_w = np.zeros((H4, W4), np.int16); tu_h = np.zeros((H4, W4), np.int16)
    cu_ix = np.zeros((H4, W4), np.int32)
    n = 0
    for i, (cx, cy, cw, ch) in enumerate(cus):
        cu_ix[cy // 4:(cy + ch) // 4, cx // 4:(cx + cw) // 4] = i
        for ty in range(cy, cy + ch, 64):
            for tx in range(cx, cx + cw, 64):
                tw, th = min(64, cw), min(64, ch)
                n += 1
                sl = (slice(ty // 4, (ty + th) // 4), slice(tx // 4, (tx + tw) // 4))
                tu_id[sl] = n; tu_w[sl] = tw; tu_h[sl] = th
    return tu_id, tu_w, tu_h, cu_ix


def gen_lf_grid(rng, cus, W, H, bit_depth=10, cu_intra=None, cu_qp=None, p_bs0=0.3):
    """LoopFilterParam rasters for a CU list, consistent with the geometry the way calcFilterStrengths
    (reference LoopFilter.cpp:495, xSetMaxFilterLengthPQFromTransformSizes :780) derives them:
    edges only at TU boundaries; luma max lengths 1/1 next to a <=4 block, else 7 (>=32) or 3 per side;
    chroma 'large' flag when both sides are >= 8 chroma samples; Bs 2 next to intra, else 1 or 0 at random."""
    ncu = len(cus)
    cu_intra = (rng.random(ncu) < 0.2) if cu_intra is None else cu_intra
    cu_qp = rng.integers(22, 45, size=ncu) if cu_qp is None else cu_qp
    tu_id, tu_w, tu_h, cu_ix = unit_maps(cus, W, H)
    H4, W4 = tu_id.shape
    out = []
    for d in (0, 1):
        g = np.zeros((H4, W4), LF_DTYPE)
        if d == 0:
            edge = np.zeros((H4, W4), bool); edge[:, 1:] = tu_id[:, 1:] != tu_id[:, :-1]
            szQ = tu_w; szP = np.zeros_like(tu_w); szP[:, 1:] = tu_w[:, :-1]
            cuP = np.zeros_like(cu_ix); cuP[:, 1:] = cu_ix[:, :-1]
        else:
            edge = np.zeros((H4, W4), bool); edge[1:, :] = tu_id[1:, :] != tu_id[:-1, :]
            szQ = tu_h; szP = np.zeros_like(tu_h); szP[1:, :] = tu_h[:-1, :]
            cuP = np.zeros_like(cu_ix); cuP[1:, :] = cu_ix[:-1, :]
        intra = cu_intra[cu_ix] | cu_intra[cuP]
        bsY = np.where(intra, 2, (rng.random((H4, W4)) >= p_bs0).astype(np.int64))
        bsU = np.where(intra, 2, (rng.random((H4, W4)) >= p_bs0).astype(np.int64))
        bsV = np.where(intra, 2, (rng.random((H4, W4)) >= p_bs0).astype(np.int64))
        bs = (bsY | (bsU << 2) | (bsV << 4)) * edge
        small = (szP <= 4) | (szQ <= 4)
        lenP = np.where(small, 1, np.where(szP >= 32, 7, 3)); lenQ = np.where(small, 1, np.where(szQ >= 32, 7, 3))
        qpl = (cu_qp[cu_ix] + cu_qp[cuP] + 1) >> 1
        g["bs"] = bs
        g["len"] = np.where(edge, 128 + (lenP << 4) + lenQ, 0)
        g["flags"] = np.where(edge & (szP >= 16) & (szQ >= 16), 32, 0) | (edge * 3)
        g["qp"][..., 0] = qpl
        g["qp"][..., 1] = np.clip(qpl - 1, 0, 63); g["qp"][..., 2] = np.clip(qpl + 1, 0, 63)
        out.append(np.ascontiguousarray(g))
    return out[0], out[1]


SAO_DTYPE = np.dtype([("type", "u1", (3,)), ("band", "u1", (3,)), ("offset", "i1", (3, 5)), ("avail", "u1"), ("rsv", "u1", (2,))])
ALFCTU_DTYPE = np.dtype([("enable", "u1", (3,)), ("lumaSet", "u1"), ("chromaAlt", "u1", (2,)), ("ccIdx", "u1", (2,))])
assert SAO_DTYPE.itemsize == 24 and ALFCTU_DTYPE.itemsize == 8
AV_L, AV_R, AV_A, AV_B, AV_AL, AV_AR, AV_BL, AV_BR = 1, 2, 4, 8, 16, 32, 64, 128


def picture_avail(ctusW, ctusH):
    """8-neighbour CTU availability when only the picture limits restrict it (single slice / tile)."""
    av = np.zeros((ctusH, ctusW), np.uint8)
    for y in range(ctusH):
        for x in range(ctusW):
            l, r, a, b = x > 0, x + 1 < ctusW, y > 0, y + 1 < ctusH
            av[y, x] = (AV_L * l | AV_R * r | AV_A * a | AV_B * b | AV_AL * (l and a) | AV_AR * (r and a)
                        | AV_BL * (l and b) | AV_BR * (r and b))
    return av.reshape(-1)


def gen_sao(rng, W, H, ctu=128, bit_depth=10, p_on=0.4, chroma=True):
    """Per-CTU SAO records (SAO on 40 % of CTUs, EO:BO 3:1, offsets in [-7,7] scaled by 1<<max(0,bd-10))."""
    ctusW, ctusH = (W + ctu - 1) // ctu, (H + ctu - 1) // ctu
    n = ctusW * ctusH
    s = np.zeros(n, SAO_DTYPE)
    s["type"] = 255
    scale = 1 << max(0, bit_depth - 10)
    for i in range(n):
        for c in range(3 if chroma else 1):
            if rng.random() < p_on:
                t = int(rng.integers(0, 4)) if rng.random() < 0.75 else 4
                s["type"][i, c] = t
                off = rng.integers(-7, 8, size=5) * scale
                if t < 4: off[2] = 0
                s["offset"][i, c] = off
                s["band"][i, c] = int(rng.integers(0, 32))
    s["avail"] = picture_avail(ctusW, ctusH)
    return s


ALF_TR = [[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12], [9, 4, 10, 8, 1, 5, 11, 7, 3, 0, 2, 6, 12],
          [0, 3, 2, 1, 8, 7, 6, 5, 4, 9, 10, 11, 12], [9, 8, 10, 4, 3, 7, 11, 5, 1, 0, 2, 6, 12]]   # AdaptiveLoopFilter.cpp:97-112


def _fixed_sets():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "alf_fixed_sets.npy"))


def gen_alf(rng, W, H, ctu=128, bit_depth=10, n_aps=2, n_chroma_alts=3, n_cc=(2, 3), p_luma=0.8, p_chroma=0.6, p_cc=0.3):
    """ALF tables (16 fixed sets + n_aps random APS sets, pre-transposed like lumaCoeffFinal) and per-CTU controls."""
    clipv = [1 << bit_depth, 1 << (bit_depth - 3), 1 << (bit_depth - 5), 1 << (bit_depth - 7)]   # m_alfClippVls
    nsets = 16 + n_aps
    coef = np.zeros((nsets, 4, 25, 13), np.int16); clip = np.zeros((nsets, 4, 25, 13), np.int16)
    coef[:16] = _fixed_sets(); clip[:16] = clipv[0]
    for s in range(16, nsets):
        base = rng.integers(-40, 41, size=(25, 13)).astype(np.int16); base[:, 12] = 128
        bclip = rng.choice(clipv, size=(25, 13)).astype(np.int16)
        for t in range(4):
            coef[s, t] = base[:, ALF_TR[t]]; clip[s, t] = bclip[:, ALF_TR[t]]
    ccoef = rng.integers(-40, 41, size=(n_chroma_alts, 7)).astype(np.int16); ccoef[:, 6] = 128
    cclip = rng.choice(clipv, size=(n_chroma_alts, 7)).astype(np.int16)
    cclip[0, :] = clipv[0]                                      # first alternative without clipping (clip index 0), as encoders mostly signal
    cc = [rng.integers(-63, 64, size=(n_cc[c], 7)).astype(np.int16) for c in range(2)]
    ctusW, ctusH = (W + ctu - 1) // ctu, (H + ctu - 1) // ctu
    n = ctusW * ctusH
    a = np.zeros(n, ALFCTU_DTYPE)
    a["enable"][:, 0] = rng.random(n) < p_luma
    a["enable"][:, 1] = rng.random(n) < p_chroma; a["enable"][:, 2] = rng.random(n) < p_chroma
    a["lumaSet"] = rng.integers(0, nsets, size=n)
    a["chromaAlt"] = rng.integers(0, n_chroma_alts, size=(n, 2))
    for c in range(2):
        a["ccIdx"][:, c] = np.where(rng.random(n) < p_cc, rng.integers(1, n_cc[c] + 1, size=n), 0) if n_cc[c] else 0
    return dict(lumaCoeff=np.ascontiguousarray(coef), lumaClip=np.ascontiguousarray(clip), chromaCoeff=ccoef, chromaClip=cclip,
                cc=cc, ctus=a)


PU_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("w", "u1"), ("h", "u1"), ("flags", "u1"), ("bcwW1", "i1"), ("refSlot", "i1", (2,)),
                     ("interDir", "u1"), ("wpIdx", "u1"), ("dmvrOff", "<u4"), ("mv", "<i4", (2, 2)), ("cpmv", "<i4", (2, 2, 2))])
assert PU_DTYPE.itemsize == 64
PU_BDOF, PU_DMVR, PU_ALTHPEL, PU_AFFINE, PU_AFFINE6, PU_PROF0, PU_PROF1, PU_GEO = 1, 2, 4, 8, 16, 32, 64, 128


def gen_pus(rng, cus, W, H, p_inter=1.0, p_bi=0.6, p_dmvr=0.35, p_bdof=0.35, p_affine=0.12, p_bcw=0.15, mv_sigma=6.0, p_int_mv=0.15, p_prof=1.0, p_geo=0.0):
    """Inter PU records for a CU list (MVs ~ N(0, 6 px) in 1/16 units; refs from 4 DPB slots:
    list 0 = {0, 1}, list 1 = {2, 3}; (0,2) and (1,3) are the equal-POC-distance pairs that allow BDOF / DMVR)."""
    recs = []
    dmvr_off = 0
    for (x, y, w, h) in cus:
        if w > 128 or h > 128 or rng.random() >= p_inter:
            continue
        r = np.zeros((), PU_DTYPE)
        r["x"], r["y"], r["w"], r["h"] = x, y, w, h
        r["bcwW1"] = 4
        mv = np.rint(rng.normal(0, mv_sigma * 16, size=(2, 2))).astype(np.int64)
        if rng.random() < p_int_mv: mv = (mv >> 4) << 4
        if rng.random() < 0.1: mv[:, rng.integers(0, 2)] &= ~15                      # one integer component
        if rng.random() < 0.03: mv += rng.integers(-3000, 3000, size=(2, 2))          # far outside the picture -> clipMv
        r["mv"] = mv
        if p_geo and 8 <= w <= 64 and 8 <= h <= 64 and w < 8 * h and h < 8 * w and rng.random() < p_geo:
            # geometric partitioning: two uni-predicted partitions (any slots, also from the same list), split direction in bcwW1
            r["refSlot"] = (int(rng.integers(0, 4)), int(rng.integers(0, 4))); r["interDir"] = 3
            r["bcwW1"] = int(rng.integers(0, 64)); r["flags"] = PU_GEO
            if w >= 8 and h >= 8 and w * h >= 128:
                r["dmvrOff"] = dmvr_off; dmvr_off += max(1, w >> 4) * max(1, h >> 4)
            recs.append(r); continue
        can_bi = (w + h) > 12
        bi = can_bi and rng.random() < p_bi
        big = w >= 8 and h >= 8 and w * h >= 128
        flags = 0
        if bi:
            u = rng.random()
            if big and u < p_dmvr:
                pair = int(rng.integers(0, 2)); r["refSlot"] = (pair, 2 + pair)
                flags |= PU_DMVR | (PU_BDOF if rng.random() < 0.8 else 0)
            elif big and u < p_dmvr + p_bdof:
                pair = int(rng.integers(0, 2)); r["refSlot"] = (pair, 2 + pair)
                flags |= PU_BDOF
            else:
                r["refSlot"] = (int(rng.integers(0, 2)), 2 + int(rng.integers(0, 2)))
                if w * h >= 256 and rng.random() < p_bcw: r["bcwW1"] = int(rng.choice([-2, 3, 5, 10]))
            r["interDir"] = 3
        else:
            l = int(rng.integers(0, 2))
            r["refSlot"] = (int(rng.integers(0, 2)), -1) if l == 0 else (-1, 2 + int(rng.integers(0, 2)))
            r["interDir"] = 1 + l
        if not (flags & (PU_DMVR | PU_BDOF)) and w >= 8 and h >= 8 and rng.random() < p_affine:
            flags |= PU_AFFINE | (PU_AFFINE6 if rng.random() < 0.5 else 0)
            if rng.random() < p_prof: flags |= PU_PROF0 | PU_PROF1   # PROF is a sequence/picture-level switch (sps_prof / ph_prof_disabled)
            big_d = rng.random() < 0.15
            for l in range(2):
                d = rng.integers(-400, 401, size=(2, 2)) if big_d else rng.integers(-24, 25, size=(2, 2))
                if rng.random() < 0.1: d[:] = 0
                r["cpmv"][l] = mv[l] + d
        elif not (flags & PU_DMVR) and rng.random() < 0.1:
            flags |= PU_ALTHPEL
        if big:
            r["dmvrOff"] = dmvr_off
            dmvr_off += max(1, w >> 4) * max(1, h >> 4)
        r["flags"] = flags
        recs.append(r)
    pus = np.array(recs, PU_DTYPE) if recs else np.zeros(0, PU_DTYPE)
    return pus, dmvr_off


WP_DTYPE = np.dtype([("w0", "<i2", (3,)), ("w1", "<i2", (3,)), ("offset", "<i2", (3,)), ("shift", "u1", (3,)), ("rsv", "u1", (3,))])
assert WP_DTYPE.itemsize == 24


def gen_wp(rng, bit_depth, pus):
    """Explicit weighted prediction for a B picture with pps_weighted_bipred: random (log2WeightDenom, iWeight, iOffset) per
    (list, refIdx, component) as a slice header carries them, the b200_wp entries WeightPrediction::getWpScaling (reference
    WeightPrediction.cpp:67-147) derives for every (refIdx0, refIdx1) combination, and pus['wpIdx'] set for the PUs they apply to
    (BcwIdx == default, InterPrediction.cpp:733).  Returns (raw int32 [2][2][3][3], entries)."""
    raw = np.zeros((2, 2, 3, 3), np.int32)
    den = [int(rng.integers(0, 8)), int(rng.integers(0, 8))]          # luma_log2_weight_denom, chroma
    for l in range(2):
        for i in range(2):
            for c in range(3):
                d = den[0] if c == 0 else den[1]
                plain = rng.random() < 0.25
                raw[l, i, c] = (d, (1 << d) if plain else (1 << d) + int(rng.integers(-128, 128)), 0 if plain else int(rng.integers(-128, 128)))
    sc = 1 << (bit_depth - 8)
    ent = np.zeros(8, WP_DTYPE); idx = {}
    k = 0
    for r0 in (-1, 0, 1):
        for r1 in (-1, 0, 1):
            if r0 < 0 and r1 < 0: continue
            e = ent[k]
            for c in range(3):
                if r0 >= 0 and r1 >= 0:
                    e["w0"][c] = raw[0, r0, c, 1]; e["w1"][c] = raw[1, r1, c, 1]
                    e["offset"][c] = (raw[0, r0, c, 2] + raw[1, r1, c, 2]) * sc; e["shift"][c] = raw[0, r0, c, 0] + 1
                else:
                    l, r = (0, r0) if r0 >= 0 else (1, r1)
                    e["w0"][c] = raw[l, r, c, 1]; e["offset"][c] = raw[l, r, c, 2] * sc; e["shift"][c] = raw[l, r, c, 0]
            idx[(r0, r1)] = k + 1; k += 1
    for p in pus:
        r0 = -1 if p["refSlot"][0] < 0 else int(p["refSlot"][0]) & 1
        r1 = -1 if p["refSlot"][1] < 0 else int(p["refSlot"][1]) & 1
        p["wpIdx"] = idx[(r0, r1)] if p["bcwW1"] == 4 and not (p["flags"] & PU_GEO) else 0    # GEO never combines with explicit weights (xPredInterBi :731)
    return raw, ent


LMCS_VPDU_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("availLeft", "u1"), ("availAbove", "u1")])


def lmcs_tables(bit_depth, min_bin, max_bin, delta_cw, chr_offset):
    """Reshape::constructReshaper (reference CommonLib/Reshape.cpp:317-373) restated for the generator: code-word deltas of the LMCS
    APS -> pivots, forward scale, chroma-scale LUT, inverse LUT.  tests/test_lmcs_oracle_vs_ref.py checks it against the real class."""
    FP = 11
    n = 1 << bit_depth; org = n // 16; l2 = org.bit_length() - 1
    bin_cw = [0] * 16
    for i in range(min_bin, max_bin + 1): bin_cw[i] = delta_cw[i] + org
    piv = [0] * 17; inp = [0] * 17; fwd = [0] * 16; inv = [0] * 16; cadj = [1 << FP] * 16
    for i in range(16):
        piv[i + 1] = piv[i] + bin_cw[i]; inp[i + 1] = inp[i] + org
        fwd[i] = (bin_cw[i] * (1 << FP) + (1 << (l2 - 1))) >> l2
        if bin_cw[i]:
            inv[i] = org * (1 << FP) // bin_cw[i]; cadj[i] = org * (1 << FP) // (bin_cw[i] + chr_offset)
    lut = np.zeros(n, np.int16)
    for v in range(n):
        idx = min_bin
        while idx <= max_bin and not v < piv[idx + 1]: idx += 1
        idx = min(idx, 15)
        lut[v] = min(max(inp[idx] + ((inv[idx] * (v - piv[idx]) + (1 << (FP - 1))) >> FP), 0), n - 1)
    return dict(orgCW=org, reshapePivot=piv, inputPivot=inp, fwdScaleCoef=fwd, chromaAdjHelpLUT=cadj, invLUT=lut)


def gen_lmcs(rng, bit_depth, cus, W, H, ctu, chroma_adj=True):
    """A legal random LMCS model (code words are multiples of 1 << (bd-5), so the pivot constraint of Reshape.cpp:355-363 holds) and the
    per-VPDU records: position of the CU covering each VPDU's top-left sample, neighbours available inside the picture (one slice, one tile)."""
    from . import abi as A
    import ctypes as C
    org = (1 << bit_depth) // 16; step = 1 << (bit_depth - 5)
    min_bin = int(rng.integers(0, 3)); max_bin = int(rng.integers(12, 16))
    while True:
        cw = rng.integers(1, 5, size=16) * step            # 0.5x .. 2x of the identity code word
        if int(cw[min_bin:max_bin + 1].sum()) <= (1 << bit_depth) - 1: break
    delta = [int(cw[i]) - org if min_bin <= i <= max_bin else 0 for i in range(16)]
    # lmcsCW[i] + lmcsDeltaCrs must stay in [OrgCW >> 3, (OrgCW << 3) - 1] (Reshape.cpp:332-333)
    chr_off = int(rng.integers(max(-7, (org >> 3) - int(cw[min_bin:max_bin + 1].min())), 8)) if chroma_adj else 0
    t = lmcs_tables(bit_depth, min_bin, max_bin, delta, chr_off)
    vs = 64 if ctu == 128 else ctu
    vW, vH = (W + vs - 1) // vs, (H + vs - 1) // vs
    # CU lookup on the 4x4 grid
    owner = np.zeros(((H + 3) // 4, (W + 3) // 4), np.int32)
    for i, (x, y, w, h) in enumerate(cus): owner[y // 4:(y + h) // 4, x // 4:(x + w) // 4] = i
    vp = np.zeros(vW * vH, LMCS_VPDU_DTYPE)
    for j in range(vH):
        for i in range(vW):
            x, y, w, h = cus[owner[j * vs // 4, i * vs // 4]]
            vp[j * vW + i] = (x, y, x > 0, y > 0)
    L = A.Lmcs(); L.chromaAdj = int(chroma_adj); L.minBinIdx = min_bin; L.maxBinIdx = max_bin; L.orgCW = t["orgCW"]
    for i in range(17): L.reshapePivot[i] = t["reshapePivot"][i]; L.inputPivot[i] = t["inputPivot"][i]
    for i in range(16): L.fwdScaleCoef[i] = t["fwdScaleCoef"][i]; L.chromaAdjHelpLUT[i] = t["chromaAdjHelpLUT"][i]
    L.invLUT = t["invLUT"].ctypes.data; L.vpdus = vp.ctypes.data
    return dict(struct=L, invLUT=t["invLUT"], vpdus=vp, minBin=min_bin, maxBin=max_bin, delta=delta, chrOff=chr_off, tables=t)


def gen_picture(rng, W, H, bit_depth=10, ctu=128, dst_slot=0, cu_kw=None, pu_kw=None, tu_kw=None, sao_p=0.4, alf_kw=None,
                deblock=True, sao=True, alf=True, lmcs=False, lmcs_chroma=True, wp=False, inter=True, given=None, cu_intra=None, intra_frac=0.0, p_ciip=0.25):
    """One synthetic post-parse picture: partition -> inter PUs (all CUs inter: intra-coded samples would
    be 'given' pixels) -> TUs/levels -> deblocking grids -> SAO / ALF CTU parameters.
    Returns a dict of numpy arrays (kept alive by the caller) plus `struct`, the abi.Picture that points into them."""
    from . import abi as A
    import ctypes as C
    if intra_frac > 0:
        # intra CUs on the device (K6): CUs in decoding order, at most 64x64 (one TU per component); a fraction of them intra, the others inter.
        # Intra CUs: TUs flagged TU_RESI (residual -> residual planes) + b200_intra_tu records with ADD_RESI where a TU carries a residual.
        cus = gen_intra_layout(rng, W, H, ctu, **({"min_size": 8} | (cu_kw or {})))
        is_intra = rng.random(len(cus)) < intra_frac
        inter_cus = [cu for cu, f in zip(cus, is_intra) if not f]; intra_cus = [cu for cu, f in zip(cus, is_intra) if f]
        pus, ndmvr = gen_pus(rng, inter_cus, W, H, **(pu_kw or {})) if inter_cus else (np.zeros(0, PU_DTYPE), 0)
        tkw = {"p_cbf": 0.35, "p_intra": 0.0, "p_lfnst": 0.0, "p_bdpcm": 0.0} | (tu_kw or {})
        tus0, coefs0 = gen_tus(rng, inter_cus, bit_depth, **tkw)
        tus1, coefs1 = gen_tus(rng, intra_cus, bit_depth, **(tkw | {"p_intra": 1.0, "p_cbf": 0.6, "p_lfnst": 0.15, "p_bdpcm": 0.05}))
        tus1["coefOff"] += len(coefs0); tus1["flags"] |= A.TU_RESI
        tus, coefs = np.concatenate([tus0, tus1]), np.concatenate([coefs0, coefs1])
        # CIIP: plain merge-like inter CUs (no BDOF / DMVR / affine / GEO / BCW, at least 64 samples) also get a planar intra block that is blended
        # with their inter prediction; the weight counts the intra CUs left (at the bottom-left corner) and above (at the top-right corner)
        cu_idx = np.full(((H + 3) // 4, (W + 3) // 4), -1, np.int64)
        for i, (x, y, w, h) in enumerate(cus): cu_idx[y // 4:(y + h) // 4, x // 4:(x + w) // 4] = i
        inter_index = [i for i, f in enumerate(is_intra) if not f]
        ciip = {}
        for j, pu in enumerate(pus):
            i = inter_index[j]; x, y, w, h = cus[i]
            if pu["flags"] == 0 and pu["bcwW1"] == 4 and w * h >= 64 and rng.random() < p_ciip:
                nl = cu_idx[(y + h - 1) // 4, x // 4 - 1] if x > 0 else -1; na = cu_idx[y // 4 - 1, (x + w - 1) // 4] if y > 0 else -1
                ciip[i] = 3 - (0 if nl >= 0 and is_intra[nl] else 1) - (0 if na >= 0 and is_intra[na] else 1)
        ciip_cus = [cus[i] for i in sorted(ciip)]
        tus2, coefs2 = gen_tus(rng, ciip_cus, bit_depth, **(tkw | {"p_cbf": 0.6}))
        tus2["coefOff"] += len(coefs0) + len(coefs1)
        for t in tus2:                                              # chroma narrower than 4 has no CIIP block (predBlendIntraCiip :891): plain inter reconstruction there
            if not (t["comp"] and (1 << int(t["log2w"])) <= 2): t["flags"] |= A.TU_RESI
        # the plain inter TUs generated above for CIIP CUs are dropped: their residual goes through K6
        ciip_pos = {(x, y) for (x, y, w, h) in ciip_cus}
        keep = np.array([((int(t["x"]) << (1 if t["comp"] else 0), int(t["y"]) << (1 if t["comp"] else 0)) not in ciip_pos) for t in tus0], bool) if len(tus0) else np.zeros(0, bool)
        tus0 = tus0[keep]
        tus, coefs = np.concatenate([tus0, tus1, tus2]), np.concatenate([coefs0, coefs1, coefs2])
        mask = is_intra.copy()
        for i in ciip: mask[i] = True
        irecs = gen_intra_records(rng, cus, W, H, only=mask, p_lm=0.2, colloc=int(rng.integers(0, 2)), ciip=ciip)
        coded = set()
        for t in np.concatenate([tus1, tus2[(tus2["flags"] & A.TU_RESI) != 0]]):
            coded.add((int(t["comp"]), int(t["x"]), int(t["y"]))); 
            if t["ict"]: coded.add((3 - int(t["comp"]), int(t["x"]), int(t["y"])))
        for r in irecs:
            if (int(r["comp"]), int(r["x"]), int(r["y"])) in coded: r["flags"] |= A.INTRA_ADD_RESI
        if cu_intra is None: cu_intra = is_intra
    else:
        cus = partition(rng, W, H, ctu=ctu, **(cu_kw or {}))
        pus, ndmvr = gen_pus(rng, cus, W, H, **(pu_kw or {})) if inter else (np.zeros(0, PU_DTYPE), 0)
        tus, coefs = gen_tus(rng, cus, bit_depth, **({"p_cbf": 0.35, "p_intra": 0.0, "p_lfnst": 0.0, "p_bdpcm": 0.0} | (tu_kw or {})))
        irecs = None
    d = dict(cus=cus, pus=pus, ndmvr=ndmvr, tus=tus, coefs=coefs)
    p = A.Picture(); p.dstSlot = dst_slot; p.flags = 0
    if irecs is not None:
        d["intraTus"] = irecs; p.intraTus = irecs.ctypes.data; p.numIntraTus = len(irecs)
    if given is not None:                                       # pre-reconstructed samples (an intra picture's predictions stand in here)
        d["given"] = given
        for c in range(3): p.given[c] = given[c].ctypes.data
    p.pus = pus.ctypes.data; p.numPus = len(pus); p.numDmvr = ndmvr + 1
    p.tus = tus.ctypes.data; p.numTus = len(tus); p.coefs = coefs.ctypes.data; p.numCoefs = len(coefs)
    if deblock:
        qp = rng.integers(22, 45, size=len(cus))
        d["lfV"], d["lfH"] = gen_lf_grid(rng, cus, W, H, bit_depth, cu_intra=np.zeros(len(cus), bool) if cu_intra is None else np.ones(len(cus), bool) if cu_intra is True else cu_intra, cu_qp=qp)
        d["lfSlices"] = np.zeros(1, LFSLICE_DTYPE)
        p.flags |= A.PIC_DEBLOCK; p.lfV = d["lfV"].ctypes.data; p.lfH = d["lfH"].ctypes.data
        p.lfSlices = d["lfSlices"].ctypes.data; p.numLfSlices = 1
    if sao:
        d["sao"] = gen_sao(rng, W, H, ctu, bit_depth, p_on=sao_p)
        p.flags |= A.PIC_SAO; p.sao = d["sao"].ctypes.data
    if alf:
        d["alf"] = gen_alf(rng, W, H, ctu, bit_depth, **(alf_kw or {}))
        d["alfTabs"] = A.make_alf_tables(d["alf"])
        p.flags |= A.PIC_ALF; p.alf = d["alf"]["ctus"].ctypes.data; p.alfTabs = C.addressof(d["alfTabs"])
    if wp:   # explicit weighted prediction: such pictures carry no BDOF / DMVR PUs (pass pu_kw=dict(p_dmvr=0, p_bdof=0))
        d["wpRaw"], d["wp"] = gen_wp(rng, bit_depth, pus)
        p.wp = d["wp"].ctypes.data; p.numWp = len(d["wp"])
    if lmcs:
        d["lmcs"] = gen_lmcs(rng, bit_depth, cus, W, H, ctu, chroma_adj=lmcs_chroma)
        p.flags |= A.PIC_LMCS; p.lmcs = C.addressof(d["lmcs"]["struct"])
    d["struct"] = p
    return d


# ---- a whole picture as a flat dict of arrays (golden fixtures): save_picture(d) -> {name: ndarray}, load_picture(z) -> the dict of gen_picture
def save_picture(d):
    out = dict(pus=d["pus"], ndmvr=np.int64(d["ndmvr"]), tus=d["tus"], coefs=d["coefs"], dstSlot=np.int64(d["struct"].dstSlot))
    for k in ("lfV", "lfH", "lfSlices", "sao", "wp", "wpRaw"):
        if k in d: out[k] = d[k]
    if "alf" in d:
        a = d["alf"]
        out.update(alf_ctus=a["ctus"], alf_lumaCoeff=a["lumaCoeff"][16:], alf_lumaClip=a["lumaClip"][16:], alf_chromaCoeff=a["chromaCoeff"],
                   alf_chromaClip=a["chromaClip"], alf_cc0=a["cc"][0], alf_cc1=a["cc"][1])
    if "lmcs" in d:
        L = d["lmcs"]["struct"]
        out.update(lmcs_scalars=np.array([L.chromaAdj, L.minBinIdx, L.maxBinIdx, L.orgCW], np.int32), lmcs_reshapePivot=np.array(list(L.reshapePivot), np.int16),
                   lmcs_inputPivot=np.array(list(L.inputPivot), np.int16), lmcs_fwdScaleCoef=np.array(list(L.fwdScaleCoef), np.int16),
                   lmcs_chromaAdjHelpLUT=np.array(list(L.chromaAdjHelpLUT), np.int32), lmcs_invLUT=d["lmcs"]["invLUT"], lmcs_vpdus=d["lmcs"]["vpdus"])
    return out


def load_picture(z, bit_depth):
    """z: mapping from save_picture (e.g. an opened .npz)."""
    fixed_sets = (_fixed_sets(), np.full((16, 4, 25, 13), 1 << bit_depth, np.int16))
    from . import abi as A
    import ctypes as C
    d = dict(pus=np.ascontiguousarray(z["pus"]), ndmvr=int(z["ndmvr"]), tus=np.ascontiguousarray(z["tus"]), coefs=np.ascontiguousarray(z["coefs"]))
    p = A.Picture(); p.dstSlot = int(z["dstSlot"]); p.flags = 0
    p.pus = d["pus"].ctypes.data; p.numPus = len(d["pus"]); p.numDmvr = d["ndmvr"] + 1
    p.tus = d["tus"].ctypes.data; p.numTus = len(d["tus"]); p.coefs = d["coefs"].ctypes.data; p.numCoefs = len(d["coefs"])
    if "lfV" in z:
        for k in ("lfV", "lfH", "lfSlices"): d[k] = np.ascontiguousarray(z[k])
        p.flags |= A.PIC_DEBLOCK; p.lfV = d["lfV"].ctypes.data; p.lfH = d["lfH"].ctypes.data; p.lfSlices = d["lfSlices"].ctypes.data; p.numLfSlices = len(d["lfSlices"])
    if "sao" in z:
        d["sao"] = np.ascontiguousarray(z["sao"]); p.flags |= A.PIC_SAO; p.sao = d["sao"].ctypes.data
    if "alf_ctus" in z:
        d["alf"] = dict(ctus=np.ascontiguousarray(z["alf_ctus"]), lumaCoeff=np.ascontiguousarray(np.concatenate([fixed_sets[0], z["alf_lumaCoeff"]])),
                        lumaClip=np.ascontiguousarray(np.concatenate([fixed_sets[1], z["alf_lumaClip"]])), chromaCoeff=np.ascontiguousarray(z["alf_chromaCoeff"]),
                        chromaClip=np.ascontiguousarray(z["alf_chromaClip"]), cc=[np.ascontiguousarray(z["alf_cc0"]), np.ascontiguousarray(z["alf_cc1"])])
        d["alfTabs"] = A.make_alf_tables(d["alf"])
        p.flags |= A.PIC_ALF; p.alf = d["alf"]["ctus"].ctypes.data; p.alfTabs = C.addressof(d["alfTabs"])
    if "wp" in z:
        d["wp"] = np.ascontiguousarray(z["wp"]); d["wpRaw"] = np.ascontiguousarray(z["wpRaw"]); p.wp = d["wp"].ctypes.data; p.numWp = len(d["wp"])
    if "lmcs_scalars" in z:
        L = A.Lmcs(); sc = z["lmcs_scalars"]; L.chromaAdj, L.minBinIdx, L.maxBinIdx, L.orgCW = [int(v) for v in sc]
        for i in range(17): L.reshapePivot[i] = int(z["lmcs_reshapePivot"][i]); L.inputPivot[i] = int(z["lmcs_inputPivot"][i])
        for i in range(16): L.fwdScaleCoef[i] = int(z["lmcs_fwdScaleCoef"][i]); L.chromaAdjHelpLUT[i] = int(z["lmcs_chromaAdjHelpLUT"][i])
        lut = np.ascontiguousarray(z["lmcs_invLUT"]); vp = np.ascontiguousarray(z["lmcs_vpdus"])
        L.invLUT = lut.ctypes.data; L.vpdus = vp.ctypes.data
        d["lmcs"] = dict(struct=L, invLUT=lut, vpdus=vp)
        p.flags |= A.PIC_LMCS; p.lmcs = C.addressof(L)
    d["struct"] = p
    return d


# ---------------------------------------------------------------------------------------------------------------- film grain
def gen_fgc_sei(rng, model_id=0, present=(1, 1, 1), max_intervals=4, max_scale=200):
    """Film grain characteristics SEI parameters in the flat int layout oracle/ref_shim.cpp:ref_film_grain reads: modelId, log2ScaleFactor, then
    per component present, numModelValues, numIntervals and per interval lower, upper, 6 model values (vvdecSEIFilmGrainCharacteristics, sei.h:208).
    Frequency-filtering model (0): values = scale, h cutoff, v cutoff (2..14); auto-regressive model (1): scale, AR coefficients."""
    out = [model_id, int(rng.integers(2, 6))]
    for c in range(3):
        if not present[c]: out += [0, 0, 0]; continue
        n = int(rng.integers(1, max_intervals + 1))
        bounds = np.sort(rng.choice(np.arange(1, 255), size=2 * n, replace=False))
        nv = 3 if model_id == 0 else 6
        out += [1, nv, n]
        for i in range(n):
            if model_id == 0: vals = [int(rng.integers(20, max_scale)), int(rng.integers(2, 15)), int(rng.integers(2, 15)), 0, 0, 0]
            else: vals = [int(rng.integers(20, max_scale)), int(rng.integers(-60, 60)), int(rng.integers(-30, 30)), int(rng.integers(-30, 30)), 1 << out[1], int(rng.integers(-20, 20))]
            out += [int(bounds[2 * i]), int(bounds[2 * i + 1])] + vals
    return np.array(out, np.int32)


def gen_film_grain_tables(rng, H):
    """Random tables in the shape FilmGrainImpl holds them (no SEI / firmware involved): for device-vs-oracle tests on the GPU box."""
    pattern = rng.integers(-127, 128, size=(2, 8, 64, 64)).astype(np.int8)
    sLUT = rng.integers(0, 256, size=(3, 256)).astype(np.uint8)
    pLUT = (rng.integers(0, 8, size=(3, 256)) << 4).astype(np.uint8)
    seeds = rng.integers(0, 1 << 32, size=(H + 15) // 16, dtype=np.uint64).astype(np.uint32)
    return pattern, sLUT, pLUT, seeds


# ---------------------------------------------------------------------------------------------------------------- intra layouts
REF_INTRA_CU_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("w", "<u2"), ("h", "<u2"), ("dirL", "u1"), ("dirC", "u1"), ("multiRefIdx", "u1"), ("bdpcm", "u1"),
                               ("bdpcmC", "u1"), ("rsv", "u1", 3)])


def gen_intra_layout(rng, W, H, ctu=128, min_size=8, max_size=64, p_split=0.75):
    """CU rectangles (luma samples) of a picture in decoding order: CTUs in raster order, inside a CTU a random quad / binary tree in z-order
    (every CU at most max_size wide and high: one TU per CU)."""
    out = []

    def rec(x, y, w, h):
        if x >= W or y >= H: return
        must = w > max_size or h > max_size or x + w > W or y + h > H
        can_q = w == h and w // 2 >= min_size
        can_v, can_h = w // 2 >= min_size, h // 2 >= min_size
        if must or ((can_q or can_v or can_h) and rng.random() < p_split):
            opts = ([0] if can_q else []) + ([] if must and (w > max_size and h > max_size) else ([1] if can_v else []) + ([2] if can_h else []))
            if not opts: opts = [0]
            k = opts[int(rng.integers(len(opts)))]
            if k == 0:
                for dy in (0, h // 2):
                    for dx in (0, w // 2): rec(x + dx, y + dy, w // 2, h // 2)
            elif k == 1: rec(x, y, w // 2, h); rec(x + w // 2, y, w // 2, h)
            else: rec(x, y, w, h // 2); rec(x, y + h // 2, w, h // 2)
        else:
            out.append((x, y, w, h))

    for cy in range(0, H, ctu):
        for cx in range(0, W, ctu): rec(cx, cy, ctu, ctu)
    return out


_INTRA_ANG = [0, 1, 2, 3, 4, 6, 8, 10, 12, 14, 16, 18, 20, 23, 26, 29, 32, 35, 39, 45, 51, 57, 64, 73, 86, 102, 128, 171, 256, 341, 512, 1024]
_INTRA_THR = [24, 24, 24, 14, 2, 0, 0, 0]


def intra_wide_angle(w, h, mode):
    """IntraPrediction::getWideAngle (IntraPrediction.cpp:443)."""
    if 1 < mode <= 66:
        shift = [0, 6, 10, 12, 14, 15][abs(int(np.log2(w)) - int(np.log2(h)))]
        if w > h and mode < 2 + shift: return mode + 65
        if h > w and mode > 66 - shift: return mode - 65
    return mode


def intra_filter_ref(w, h, mode, mrl, bdpcm):
    """IntraPrediction::useFilteredIntraRefSamples (:1301) for a luma block."""
    if mrl or bdpcm or mode == 1: return False
    if mode == 0: return w * h > 32
    pm = intra_wide_angle(w, h, mode)
    diff = min(abs(pm - 18), abs(pm - 50))
    ang = _INTRA_ANG[abs(pm - 50 if pm >= 34 else -(pm - 18))]
    return diff > _INTRA_THR[(int(np.log2(w)) + int(np.log2(h))) >> 1] and (ang & 31) == 0


def isp_regions(w, h, split):
    """Prediction regions (dx, dy, rw, rh) of an ISP CU and the width of its transform units (CU::getISPSplitDim UnitTools.cpp:360, isPredRegDiffFromTB):
    split 1 = horizontal (regions stacked), 2 = vertical; vertical sub-partitions narrower than 4 share a 4-wide region."""
    sd, nd = (h, w) if split == 1 else (w, h)
    part = max(sd >> 2, 16 // nd if nd < 16 else 1)
    if split == 1: return [(0, k * part, w, part) for k in range(h // part)], w
    rw = max(part, 4)
    return [(k * rw, 0, rw, h) for k in range(w // rw)], part


def gen_intra_records(rng, layout, W, H, modes=None, p_mrl=0.15, p_bdpcm=0.08, p_resi=0.0, upto=None, only=None, p_mip=0.15, colloc=0, p_lm=0.0, ciip=None, p_isp=0.0):
    """b200_intra_tu records (Y, Cb, Cr per CU, decoding order) for a single-tree all-intra layout of gen_intra_layout: random modes, MRL on some luma
    blocks, BDPCM prediction on some, availability as xFillReferenceSamples derives it from the decoding order (pinned against the reference's own
    analysis through the glue flattener by tests/test_intra_oracle_vs_ref.py).  Luma blocks whose chroma would be narrower than 4 or smaller than 16
    samples are luma-only (local dual tree; their chroma is not generated)."""
    A = abi
    owner = np.full(((H + 3) // 4, (W + 3) // 4), 1 << 30, np.int64)
    for i, (x, y, w, h) in enumerate(layout): owner[y // 4:(y + h) // 4, x // 4:(x + w) // 4] = i
    recs = []
    chroma_modes = [0, 1, 18, 50, 2, 34, 66, -1, -1, -1, 23, 45, 61]
    for i, (x, y, w, h) in enumerate(layout if upto is None else layout[:upto + 1]):
        mip = 0
        if modes is not None and i in modes:
            m = modes[i]; dirL, dirC, mrl, bdpcm = m[:4]; mip = m[4] if len(m) > 4 else 0
        else:
            dirL, mrl, bdpcm = int(rng.integers(0, 67)), 0, 0
            if rng.random() < p_mrl and y % 128: mrl, dirL = int(rng.integers(1, 3)), int(rng.integers(1, 67))
            elif rng.random() < p_bdpcm and w <= 32 and h <= 32: bdpcm = int(rng.integers(1, 3))
            elif rng.random() < p_mip:                                   # matrix intra prediction: dirL is the MIP mode index of the size class
                n_modes = 16 if (w, h) == (4, 4) else 8 if (w == 4 or h == 4 or (w, h) == (8, 8)) else 6
                dirL, mip = int(rng.integers(0, n_modes)), 1 | (int(rng.integers(0, 2)) << 1)
            dirC = chroma_modes[int(rng.integers(len(chroma_modes)))]
            if rng.random() < p_lm: dirC = int(rng.integers(67, 70))
        if dirC < 0 or dirC == 70: dirC = 0 if mip else dirL            # DM (a MIP luma CU counts as planar: PU::getCoLocatedIntraLumaMode)
        wc = ciip.get(i, 0) if ciip else 0                              # CIIP CU (inter): planar blocks blended with the inter prediction, weight wc
        if wc: dirL, dirC, mrl, bdpcm, mip = 0, 0, 0, 0, 0
        lm = dirC in (67, 68, 69)                                       # LM_CHROMA_IDX, MDLM_L_IDX, MDLM_T_IDX
        if only is not None and not only[i]: continue                   # an inter CU of a mixed picture: a neighbour, not a block of the list
        def avail(ux, uy): return 0 <= ux < owner.shape[1] and 0 <= uy < owner.shape[0] and owner[uy, ux] < i
        tl = avail(x // 4 - 1, y // 4 - 1)
        na = 0
        while na < 2 * w // 4 and avail(x // 4 + na, y // 4 - 1): na += 1
        nl = 0
        while nl < 2 * h // 4 and avail(x // 4 - 1, y // 4 + nl): nl += 1
        luma_only = w < 8 or (w // 2) * (h // 2) < 16
        if wc: luma_only = (w // 2) <= 2
        isp = 0
        if p_isp and not (mip or mrl or bdpcm or wc) and w * h > 16 and rng.random() < p_isp: isp = int(rng.integers(1, 3))
        if isp:
            # intra sub-partitions: one record per prediction region, the CU-level neighbourhood in each (include/vvdec_b200.h, B200_INTRA_ISP)
            regions, tuw = isp_regions(w, h, isp)
            for k, (dx, dy, rw, rh) in enumerate(regions):
                r = np.zeros((), A.INTRA_TU_DTYPE)
                r["x"], r["y"], r["log2w"], r["log2h"], r["mode"] = x + dx, y + dy, int(np.log2(rw)), int(np.log2(rh)), dirL
                r["mip"] = isp | (k << 2) | (int(np.log2(len(regions))) << 4)
                mask = int(rng.integers(0, 1 << (rw // tuw))) if rng.random() < max(p_resi, 0.0) * 1.5 else 0
                r["ciip"] = mask
                r["flags"] = A.INTRA_ISP | (A.INTRA_AVAIL_TL if tl else 0) | (4 if mask else 0)
                r["numAbove"], r["numLeft"], r["lmLeft"], r["lmAbove"] = na, nl, int(avail(x // 4 - 1, y // 4)), int(avail(x // 4, y // 4 - 1))
                recs.append(r)
        for c in range(1 if luma_only else 3):
            if isp and c == 0: continue
            r = np.zeros((), A.INTRA_TU_DTYPE)
            r["ciip"] = wc
            sh = 1 if c else 0
            r["x"], r["y"], r["log2w"], r["log2h"], r["comp"] = x >> sh, y >> sh, int(np.log2(w >> sh)), int(np.log2(h >> sh)), c
            if c == 0 and mip: r["mode"], r["mip"] = A.INTRA_MIP, dirL | ((mip >> 1) << 7)
            elif c == 0: r["mode"] = (A.INTRA_BDPCM_HOR if bdpcm == 1 else A.INTRA_BDPCM_VER) if bdpcm else dirL
            else: r["mode"] = dirC + 3 if lm else dirC                # B200_INTRA_LM = 70 ...
            r["multiRefIdx"] = mrl if c == 0 else 0
            fl = A.INTRA_AVAIL_TL if tl else 0
            if c == 0 and not mip and intra_filter_ref(w, h, dirL, mrl, bdpcm): fl |= A.INTRA_FILTER_REF
            if rng.random() < p_resi: fl |= 4
            if c and lm:
                cw, chh = w >> 1, h >> 1
                if na: fl |= A.INTRA_LM_ABOVE
                if nl: fl |= A.INTRA_LM_LEFT
                if colloc: fl |= A.INTRA_LM_COLLOCATED
                r["lmAbove"] = cw // 2 if na else 0; r["lmLeft"] = chh // 2 if nl else 0
                if dirC == 69 and na: r["lmAbove"] = min(na, cw // 2 + min(cw // 2, chh // 2))
                if dirC == 68 and nl: r["lmLeft"] = min(nl, chh // 2 + min(chh // 2, cw // 2))
            r["flags"], r["numAbove"], r["numLeft"] = fl, na, nl
            recs.append(r)
    return np.array(recs, A.INTRA_TU_DTYPE)
