"""BASELINE.json configs on the GPU, at their full sizes, against the pinned oracle (bit-exact):
config[1] 1080p RA picture chain           -> tests/test_picture_gpu.py::test_gop_decompress[2-1920-1080-...]
config[2] 3840x2160 RA, full in-loop chain  -> test_config2_4k_full_chain
config[3] 3840x2160 DMVR+BDOF+PROF-heavy MC -> test_config3_4k_refinement_heavy_mc
config[4] 7680x4320 RA (one GOP's picture; the 8-GOP sharding itself is tests/test_gop_shard_cpu.py) -> test_config4_8k_picture"""
import ctypes as C
import numpy as np
import pytest
import vvdec_b200
from vvdec_b200 import abi, synth
from tests.helpers import ref_ptrs, oracle_decompress
from tests.test_k2_oracle_vs_ref import _case

pytestmark = pytest.mark.gpu


def _picture(b200, oracle, W, H, seed, **kw):
    rng = np.random.default_rng(seed)
    bd = 10
    g = abi.make_geom(W, H, bd)
    ctx = C.c_void_p()
    vvdec_b200.check(b200.b200_ctx_create(C.byref(ctx), C.byref(g), 5, 1, -1))
    try:
        dpb = [synth.noise_planes(rng, W, H, bd) for _ in range(4)]
        for s in range(4): vvdec_b200.check(b200.b200_ctx_load_slot(ctx, s, abi.plane_ptrs(dpb[s])))
        pic = synth.gen_picture(rng, W, H, bd, dst_slot=4, **kw)
        want, dm_want = oracle_decompress(oracle, g, dpb, pic)
        h = b200.b200_decompress_picture(ctx, C.byref(pic["struct"])); assert h >= 0, b200.b200_last_error()
        dm = np.zeros((pic["ndmvr"] + 1, 2), np.int32)
        vvdec_b200.check(b200.b200_wait_picture(ctx, h, dm.ctypes.data, len(dm)))
        got = [np.zeros_like(p) for p in want]
        vvdec_b200.check(b200.b200_get_frame(ctx, 4, abi.plane_ptrs(got)))
        for c in range(3):
            assert np.array_equal(want[c], got[c]), f"plane {c}: {len(np.argwhere(want[c] != got[c]))} diffs"
        assert np.array_equal(dm, dm_want)
    finally:
        b200.b200_ctx_destroy(ctx)


def test_config2_4k_full_chain(b200, oracle):
    _picture(b200, oracle, 3840, 2160, 42)


def test_config3_4k_refinement_heavy_mc(b200, oracle):
    """DMVR + BDOF + PROF-heavy inter content: 90 % bi-prediction, 45 % DMVR, 40 % BDOF, 15 % affine with PROF."""
    W, H, bd = 3840, 2160, 10
    pus, nd, refs = _case(43, W, H, bd, p_bi=0.9, p_dmvr=0.45, p_bdof=0.40, p_affine=0.15, p_prof=1.0, mv_sigma=3.0)
    fl = pus["flags"]; area = pus["w"].astype(int) * pus["h"]
    # share of the picture area: a third goes through DMVR, more than half through BDOF (small PUs cannot have either)
    assert area[(fl & 2) != 0].sum() > 0.3 * area.sum() and area[(fl & 1) != 0].sum() > 0.5 * area.sum() and ((fl & 32) != 0).any()
    g = abi.make_geom(W, H, bd)
    a = [np.full((H, W), -1, np.int16), np.full((H // 2, W // 2), -1, np.int16), np.full((H // 2, W // 2), -1, np.int16)]
    b = [p.copy() for p in a]
    da = np.zeros((nd + 1, 2), np.int32); db = np.zeros((nd + 1, 2), np.int32)
    rp = ref_ptrs(refs)
    oracle.orc_mc_predict(C.byref(g), abi.plane_ptrs(a), rp, pus.ctypes.data, len(pus), da.ctypes.data)
    vvdec_b200.check(b200.b200_mc_predict(C.byref(g), abi.plane_ptrs(b), rp, 4, pus.ctypes.data, len(pus), db.ctypes.data, nd + 1))
    for c in range(3): assert np.array_equal(a[c], b[c]), f"plane {c}"
    assert np.array_equal(da, db)


def test_config4_8k_picture(b200, oracle):
    _picture(b200, oracle, 7680, 4320, 44)
