"""The inter flattener (vvdec_b200/vvdec_glue/flatten_pu.h) against the reference: real CodingUnits (merge / MMVD / SMVD / BCW / IMV /
affine / SbTMVP syntax) -> the real InterPrediction::motionCompensation on one hand, flattenPU -> the pinned oracle on the other.
The flattener takes the reference's decisions (BDOF, DMVR, identical-motion fallback, weighted prediction, SbTMVP run merging) with
the reference's own predicates, so both sides must produce the same samples and DMVR deltas."""
import ctypes as C
import numpy as np
import pytest
from vvdec_b200 import abi, synth
from tests.helpers import ref_ptrs

pytestmark = pytest.mark.ref


class CuSyntax(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32), ("refIdx", C.c_int32 * 2), ("mv", C.c_int32 * 2 * 3 * 2),
                ("affine", C.c_int32), ("affine6", C.c_int32), ("mergeFlag", C.c_int32), ("mmvdFlag", C.c_int32), ("smvd", C.c_int32),
                ("bcwIdx", C.c_int32), ("imvHpel", C.c_int32), ("sbTmvp", C.c_int32), ("sbSeed", C.c_int32),
                ("geo", C.c_int32), ("geoSplitDir", C.c_int32), ("geoDir0", C.c_int32), ("geoDir1", C.c_int32)]


def gen_cus(rng, W, H, wp):
    cus = synth.partition(rng, W, H)
    out = []
    for (x, y, w, h) in cus:
        if w > 128 or h > 128: continue
        c = CuSyntax(); c.x, c.y, c.w, c.h = int(x), int(y), int(w), int(h)
        bi = (w + h) > 12 and rng.random() < 0.65
        if bi: c.refIdx[0], c.refIdx[1] = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        elif rng.random() < 0.5: c.refIdx[0], c.refIdx[1] = int(rng.integers(0, 2)), -1
        else: c.refIdx[0], c.refIdx[1] = -1, int(rng.integers(0, 2))
        base = np.rint(rng.normal(0, 6 * 16, size=(2, 2))).astype(int)
        if rng.random() < 0.15: base[1] = base[0]                       # same motion in both lists (identical-motion fallback with altRefs)
        if rng.random() < 0.03: base += rng.integers(-3000, 3000, size=(2, 2))
        for l in range(2):
            for k in range(3):
                d = rng.integers(-24, 25, size=2) if k else np.zeros(2, int)
                c.mv[l][k][0], c.mv[l][k][1] = int(base[l][0] + d[0]), int(base[l][1] + d[1])
        u = rng.random()
        if 8 <= w <= 64 and 8 <= h <= 64 and w < 8 * h and h < 8 * w and rng.random() < 0.12:
            c.geo = 1; c.geoSplitDir = int(rng.integers(0, 64)); c.mergeFlag = 1
            c.geoDir0 = ((int(rng.integers(0, 2)) + 1) << 4) | int(rng.integers(0, 2)); c.geoDir1 = ((int(rng.integers(0, 2)) + 1) << 4) | int(rng.integers(0, 2))
        elif w >= 8 and h >= 8 and u < 0.15:
            c.affine = 1; c.affine6 = int(rng.random() < 0.5)
            if rng.random() < 0.2:
                for k in range(3): c.mv[1][k][0], c.mv[1][k][1] = c.mv[0][k][0], c.mv[0][k][1]
        elif w >= 8 and h >= 8 and u < 0.25:
            c.sbTmvp = 1; c.sbSeed = int(rng.integers(1, 1 << 30)); c.mergeFlag = 1
        else:
            c.mergeFlag = int(rng.random() < 0.6); c.mmvdFlag = int(c.mergeFlag and rng.random() < 0.2)
            c.smvd = int(bi and not c.mergeFlag and rng.random() < 0.2)
            c.imvHpel = int(not c.mergeFlag and rng.random() < 0.15)
            if c.imvHpel:                                               # half-sample AMVR: MVs are multiples of 8
                for l in range(2): c.mv[l][0][0] &= ~7; c.mv[l][0][1] &= ~7
        if bi and not c.sbTmvp and not c.geo and w * h >= 256 and rng.random() < 0.2: c.bcwIdx = int(rng.integers(1, 5))   # internal-domain index, BCW_DEFAULT = 0
        else: c.bcwIdx = 0
        out.append(c)
    return (CuSyntax * len(out))(*out)


@pytest.mark.parametrize("simd", [0, 1])
@pytest.mark.parametrize("alt_refs,wp", [(0, False), (1, False), (0, True)])
def test_flatten_pu(oracle, ref, simd, alt_refs, wp):
    W, H, bd = 384, 256, 10
    for seed in (1, 2, 3):
        rng = np.random.default_rng(seed * 10 + alt_refs + 2 * wp)
        g = abi.make_geom(W, H, bd)
        refs = [synth.noise_planes(rng, W, H, bd) for _ in range(4)]
        slot_planes = refs if not alt_refs else refs                    # slots are physical pictures; altRefs only changes list 1's mapping
        cus = gen_cus(rng, W, H, wp)
        raw = ent = None
        if wp:
            raw, ent = synth.gen_wp(rng, bd, np.zeros(0, synth.PU_DTYPE))
            ref.ref_set_wp(raw.ctypes.data)
        want = [np.zeros((H, W), np.int16), np.zeros((H // 2, W // 2), np.int16), np.zeros((H // 2, W // 2), np.int16)]
        recs = np.zeros(len(cus) * 40, synth.PU_DTYPE)
        nd = 4096
        dm_want = np.zeros((nd, 2), np.int32)
        try:
            n = ref.ref_flatten_pu_case(simd, C.byref(g), ref_ptrs(slot_planes), alt_refs, cus, len(cus), abi.plane_ptrs(want), recs.ctypes.data, len(recs), dm_want.ctypes.data, nd)
        finally:
            ref.ref_set_wp(None)
        assert 0 < n <= len(recs), n
        recs = recs[:n]
        got = [np.zeros_like(p) for p in want]
        dm = np.zeros((nd, 2), np.int32)
        oracle.orc_mc_predict_wp(C.byref(g), abi.plane_ptrs(got), ref_ptrs(slot_planes), recs.ctypes.data, n, dm.ctypes.data, ent.ctypes.data if wp else None)
        for c in range(3):
            if not np.array_equal(want[c], got[c]):
                d = np.argwhere(want[c] != got[c]); y, x = d[0]; sh = 1 if c else 0
                hit = [p for p in recs if p["x"] >> sh <= x < (p["x"] + p["w"]) >> sh and p["y"] >> sh <= y < (p["y"] + p["h"]) >> sh]
                raise AssertionError(f"seed {seed} plane {c}: {len(d)} diffs, first at {(y, x)}; record {hit[:1]}")
        assert np.array_equal(dm, dm_want)
        kinds = recs["flags"]
        assert (kinds & 128).any() and (kinds & 8).any() and (wp or ((kinds & 1).any() and (kinds & 2).any()))   # affine, and (without explicit weights) BDOF and DMVR occurred
        if wp: assert (recs["wpIdx"] != 0).any()
