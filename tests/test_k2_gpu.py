"""K2 parity on the GPU: CUDA inter prediction (through the C ABI) vs the pinned oracle, bit-exact (samples and DMVR MV deltas)."""
import ctypes as C
import numpy as np
import pytest
import vvdec_b200
from vvdec_b200 import abi, synth
from tests.helpers import ref_ptrs
from tests.test_k2_oracle_vs_ref import _case

pytestmark = pytest.mark.gpu


def _compare(b200, oracle, W, H, bd, pus, ndmvr, refs):
    g = abi.make_geom(W, H, bd)
    a = [np.full((H, W), -1, np.int16), np.full((H // 2, W // 2), -1, np.int16), np.full((H // 2, W // 2), -1, np.int16)]
    b = [p.copy() for p in a]
    da = np.zeros((ndmvr + 1, 2), np.int32); db = np.zeros((ndmvr + 1, 2), np.int32)
    rp = ref_ptrs(refs)
    oracle.orc_mc_predict(C.byref(g), abi.plane_ptrs(a), rp, pus.ctypes.data, len(pus), da.ctypes.data)
    vvdec_b200.check(b200.b200_mc_predict(C.byref(g), abi.plane_ptrs(b), rp, 4, pus.ctypes.data, len(pus), db.ctypes.data, ndmvr + 1))
    for c in range(3):
        if not np.array_equal(a[c], b[c]):
            d = np.argwhere(a[c] != b[c]); y, x = d[0]; sh = 1 if c else 0
            hit = [i for i, p in enumerate(pus) if p["x"] >> sh <= x < (p["x"] + p["w"]) >> sh and p["y"] >> sh <= y < (p["y"] + p["h"]) >> sh]
            raise AssertionError(f"plane {c}: {len(d)} diffs, first at {(y, x)}: {a[c][y, x]} vs {b[c][y, x]}; PU {pus[hit[0]] if hit else None}")
    assert np.array_equal(da, db), f"DMVR deltas differ: {np.argwhere(da != db)[:5]}"


@pytest.mark.parametrize("name,kw", [("regular", dict(p_dmvr=0, p_bdof=0, p_affine=0)), ("bdof", dict(p_dmvr=0, p_bdof=0.9, p_affine=0, p_bi=0.9)),
                                      ("dmvr", dict(p_dmvr=0.9, p_bdof=0.05, p_affine=0, p_bi=0.9, mv_sigma=2.0)),
                                      ("affine", dict(p_dmvr=0, p_bdof=0, p_affine=0.9, p_prof=0.8)),
                                      ("geo", dict(p_geo=0.7, p_dmvr=0.1, p_bdof=0.1))])
def test_mc_modes(b200, oracle, name, kw):
    pus, nd, refs = _case(11, 416, 240, 10, **kw)
    _compare(b200, oracle, 416, 240, 10, pus, nd, refs)


@pytest.mark.parametrize("seed,W,H,bd", [(5, 1920, 1080, 10), (6, 256, 128, 8), (7, 384, 256, 12), (8, 3840, 2160, 10)])
def test_mc_mixed_pictures(b200, oracle, seed, W, H, bd):
    pus, nd, refs = _case(seed, W, H, bd, **({"p_dmvr": 0.0} if bd > 10 else {}))
    _compare(b200, oracle, W, H, bd, pus, nd, refs)


@pytest.mark.parametrize("seed,W,H,bd", [(21, 416, 240, 10), (22, 256, 128, 8), (23, 1920, 1080, 10)])
def test_mc_explicit_weighted_prediction(b200, oracle, seed, W, H, bd):
    """b200_mc_predict_wp: uni / bi / affine(+PROF) PUs with explicit weights, BCW PUs that bypass them (wpIdx 0)."""
    pus, nd, refs = _case(seed, W, H, bd, p_dmvr=0.0, p_bdof=0.0, p_affine=0.2, p_bcw=0.3)
    raw, ent = synth.gen_wp(np.random.default_rng(seed), bd, pus)
    g = abi.make_geom(W, H, bd)
    a = [np.full((H, W), -1, np.int16), np.full((H // 2, W // 2), -1, np.int16), np.full((H // 2, W // 2), -1, np.int16)]
    b = [p.copy() for p in a]
    da = np.zeros((nd + 1, 2), np.int32); db = np.zeros((nd + 1, 2), np.int32)
    rp = ref_ptrs(refs)
    oracle.orc_mc_predict_wp(C.byref(g), abi.plane_ptrs(a), rp, pus.ctypes.data, len(pus), da.ctypes.data, ent.ctypes.data)
    vvdec_b200.check(b200.b200_mc_predict_wp(C.byref(g), abi.plane_ptrs(b), rp, 4, pus.ctypes.data, len(pus), db.ctypes.data, nd + 1, ent.ctypes.data, len(ent)))
    for c in range(3):
        assert np.array_equal(a[c], b[c]), f"plane {c}: {len(np.argwhere(a[c] != b[c]))} diffs"
    # a PU that asks for weights together with BDOF is refused
    bad = pus.copy(); i = int(np.argmax((bad["refSlot"][:, 0] >= 0) & (bad["refSlot"][:, 1] >= 0) & (bad["w"] >= 8) & (bad["h"] >= 16) & ((bad["flags"] & 8) == 0)))
    bad["flags"][i] |= 1; bad["wpIdx"][i] = 1; bad["bcwW1"][i] = 4
    assert b200.b200_mc_predict_wp(C.byref(g), abi.plane_ptrs(b), rp, 4, bad.ctypes.data, len(bad), db.ctypes.data, nd + 1, ent.ctypes.data, len(ent)) == -2
