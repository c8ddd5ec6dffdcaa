"""Picture-level parity on the GPU: b200_decompress_picture (DecLibRecon seam, device-resident DPB, all five kernel families
chained) vs the pinned oracle chain, over a short GOP where later pictures reference earlier reconstructed ones."""
import ctypes as C
import numpy as np
import pytest
import vvdec_b200
from vvdec_b200 import abi, synth
from tests.helpers import oracle_decompress

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,W,H,flagsets", [(1, 416, 240, [(1, 1, 1), (1, 0, 1), (0, 0, 0), (1, 1, 0)]), (2, 1920, 1080, [(1, 1, 1), (1, 1, 1)])])
def test_gop_decompress(b200, oracle, seed, W, H, flagsets):
    rng = np.random.default_rng(seed)
    bd = 10
    g = abi.make_geom(W, H, bd)
    ctx = C.c_void_p()
    vvdec_b200.check(b200.b200_ctx_create(C.byref(ctx), C.byref(g), 6, 3, -1))
    try:
        dpb = [synth.noise_planes(rng, W, H, bd) for _ in range(4)] + [None, None]
        for s in range(4):
            vvdec_b200.check(b200.b200_ctx_load_slot(ctx, s, abi.plane_ptrs(dpb[s])))
        dpb[4] = [p.copy() for p in dpb[0]]; dpb[5] = [p.copy() for p in dpb[0]]
        order = [4, 5, 0, 2, 1, 3]          # destination slots: later pictures overwrite the initial references
        for i, (db, sa, al) in enumerate(flagsets):
            dst = order[i % len(order)]
            pic = synth.gen_picture(rng, W, H, bd, dst_slot=dst, deblock=db, sao=sa, alf=al)
            want, dm_want = oracle_decompress(oracle, g, dpb[:4], pic)
            h = b200.b200_decompress_picture(ctx, C.byref(pic["struct"]))
            assert h >= 0, b200.b200_last_error()
            dm = np.zeros((pic["ndmvr"] + 1, 2), np.int32)
            vvdec_b200.check(b200.b200_wait_picture(ctx, h, dm.ctypes.data, len(dm)))
            got = [np.zeros_like(p) for p in want]
            vvdec_b200.check(b200.b200_get_frame(ctx, dst, abi.plane_ptrs(got)))
            for c in range(3):
                assert np.array_equal(want[c], got[c]), f"picture {i} plane {c}: {len(np.argwhere(want[c] != got[c]))} diffs"
            assert np.array_equal(dm, dm_want)
            dpb[dst] = want
        assert b200.b200_ctx_kernel_launches(ctx) >= 5
    finally:
        b200.b200_ctx_destroy(ctx)


def test_ctx_errors(b200):
    g = abi.make_geom(100, 64, 10)
    ctx = C.c_void_p()
    assert b200.b200_ctx_create(C.byref(ctx), C.byref(g), 4, 2, -1) == -2
    assert b"multiple of 8" in b200.b200_last_error()


def test_invalid_records_are_reported(b200):
    """Records are validated on the device while they are bucketed: a PU pointing at a DPB slot the context does not have (or a TU with
    an impossible size) makes b200_pic_run / b200_decompress_picture refuse the picture with B200_ERR_PARAM."""
    rng = np.random.default_rng(5)
    W, H, bd = 416, 240, 10
    g = abi.make_geom(W, H, bd)
    ctx = C.c_void_p()
    vvdec_b200.check(b200.b200_ctx_create(C.byref(ctx), C.byref(g), 6, 2, -1))
    try:
        ref = synth.noise_planes(rng, W, H, bd)
        for s in range(6): vvdec_b200.check(b200.b200_ctx_load_slot(ctx, s, abi.plane_ptrs(ref)))
        pic = synth.gen_picture(rng, W, H, bd, dst_slot=4, deblock=0, sao=0, alf=0)
        h = b200.b200_decompress_picture(ctx, C.byref(pic["struct"])); assert h >= 0
        assert b200.b200_wait_picture(ctx, h, None, 0) == 0
        orig = pic["pus"]["refSlot"][3].copy()
        pic["pus"]["refSlot"][3] = (17, -1)
        assert b200.b200_decompress_picture(ctx, C.byref(pic["struct"])) == -2 and b"PU list" in b200.b200_last_error()
        pic["pus"]["refSlot"][3] = orig
        pic["tus"]["log2w"][0] = 7
        assert b200.b200_decompress_picture(ctx, C.byref(pic["struct"])) == -2 and b"TU list" in b200.b200_last_error()
        pic["tus"]["log2w"][0] = 2
        # records that would touch memory outside the picture or outside the uploaded arrays are refused too
        for arr, field, idx, bad, what in (("pus", "x", 5, W - 4 + 8, b"PU list"), ("pus", "y", 5, H, b"PU list"), ("pus", "x", 5, 2, b"PU list"),
                                           ("tus", "coefOff", 1, 1 << 30, b"TU list"), ("tus", "x", 1, W, b"TU list"), ("tus", "maxX", 1, 200, b"TU list")):
            keep = pic[arr][field][idx].copy()
            pic[arr][field][idx] = bad
            assert b200.b200_decompress_picture(ctx, C.byref(pic["struct"])) == -2 and what in b200.b200_last_error(), (arr, field)
            pic[arr][field][idx] = keep
        # fields the kernels index tables / working sets with: LFNST byte, joint-CbCr mode on a luma TU, transform-skip / BDPCM blocks wider than 32
        tus = pic["tus"]
        i_luma = int(np.flatnonzero((tus["comp"] == 0) & (tus["log2w"] >= 2) & (tus["log2h"] >= 2) & ((tus["flags"] & 7) == 0))[0])
        for field, bad in (("lfnst", 0x10), ("lfnst", 3), ("lfnst", 0x41), ("ict", 2)):
            keep = tus[field][i_luma].copy(); tus[field][i_luma] = bad
            assert b200.b200_decompress_picture(ctx, C.byref(pic["struct"])) == -2 and b"TU list" in b200.b200_last_error(), (field, bad)
            tus[field][i_luma] = keep
        keep = tus[i_luma].copy()
        tus["flags"][i_luma] |= 1; tus["log2w"][i_luma] = 6; tus["x"][i_luma] = 0                      # a 64-wide transform-skip block
        assert b200.b200_decompress_picture(ctx, C.byref(pic["struct"])) == -2 and b"TU list" in b200.b200_last_error()
        tus[i_luma] = keep
        tus["flags"][i_luma] |= 2                                                                      # BDPCM without transform skip / with a partial corner
        assert b200.b200_decompress_picture(ctx, C.byref(pic["struct"])) == -2 and b"TU list" in b200.b200_last_error()
        tus[i_luma] = keep
        fpic = synth.gen_picture(rng, W, H, bd, dst_slot=4)               # with filters: CTU records are range-checked too
        for arr, field, bad in (("alf", "lumaSet", 250), ("sao", "type", 7), ("alf", "ccIdx", 200)):
            recs = fpic["alf"]["ctus"] if arr == "alf" else fpic["sao"]
            keep = recs[0].copy()
            recs[field][0] = bad
            if arr == "alf": recs["enable"][0] = 1
            assert b200.b200_decompress_picture(ctx, C.byref(fpic["struct"])) == -2 and b"CTU record" in b200.b200_last_error(), (arr, field)
            recs[0] = keep
        h = b200.b200_decompress_picture(ctx, C.byref(fpic["struct"])); assert h >= 0
        assert b200.b200_wait_picture(ctx, h, None, 0) == 0
        dm = np.flatnonzero(pic["pus"]["flags"] & synth.PU_DMVR)
        if dm.size:
            keep = pic["pus"]["dmvrOff"][dm[0]].copy()
            pic["pus"]["dmvrOff"][dm[0]] = 1 << 30
            assert b200.b200_decompress_picture(ctx, C.byref(pic["struct"])) == -2 and b"PU list" in b200.b200_last_error()
            pic["pus"]["dmvrOff"][dm[0]] = keep
        h = b200.b200_decompress_picture(ctx, C.byref(pic["struct"])); assert h >= 0      # the context is still usable
        assert b200.b200_wait_picture(ctx, h, None, 0) == 0
    finally:
        b200.b200_ctx_destroy(ctx)


@pytest.mark.parametrize("W,H,ctu,chroma_adj", [(416, 240, 128, True), (416, 240, 64, True), (832, 480, 128, False), (1920, 1080, 128, True)])
def test_lmcs_picture(b200, oracle, W, H, ctu, chroma_adj):
    """LMCS on: forward-mapped luma prediction fused into K2, luma TUs -> per-VPDU chroma scale -> scaled chroma
    TUs, inverse map before deblocking — against the oracle chain (which tests/test_lmcs_oracle_vs_ref.py pins to the real Reshape)."""
    rng = np.random.default_rng(W + ctu)
    bd = 10
    g = abi.make_geom(W, H, bd, ctu=ctu)
    ctx = C.c_void_p()
    vvdec_b200.check(b200.b200_ctx_create(C.byref(ctx), C.byref(g), 6, 2, -1))
    try:
        dpb = [synth.noise_planes(rng, W, H, bd) for _ in range(4)]
        for s in range(4): vvdec_b200.check(b200.b200_ctx_load_slot(ctx, s, abi.plane_ptrs(dpb[s])))
        for k in range(2):
            pic = synth.gen_picture(rng, W, H, bd, ctu=ctu, dst_slot=4 + k, lmcs=True, lmcs_chroma=chroma_adj)
            want, dm_want = oracle_decompress(oracle, g, dpb, pic)
            h = b200.b200_decompress_picture(ctx, C.byref(pic["struct"]))
            assert h >= 0, b200.b200_last_error()
            dm = np.zeros((pic["ndmvr"] + 1, 2), np.int32)
            vvdec_b200.check(b200.b200_wait_picture(ctx, h, dm.ctypes.data, len(dm)))
            got = [np.zeros_like(p) for p in want]
            vvdec_b200.check(b200.b200_get_frame(ctx, 4 + k, abi.plane_ptrs(got)))
            for c in range(3):
                assert np.array_equal(want[c], got[c]), f"picture {k} plane {c}: {len(np.argwhere(want[c] != got[c]))} diffs"
            assert np.array_equal(dm, dm_want)
    finally:
        b200.b200_ctx_destroy(ctx)


def test_weighted_prediction_picture(b200, oracle):
    """Explicit weighted prediction and GEO at picture level (together with LMCS: the combined luma is what gets forward-mapped)."""
    W, H, bd = 832, 480, 10
    rng = np.random.default_rng(31)
    g = abi.make_geom(W, H, bd)
    ctx = C.c_void_p()
    vvdec_b200.check(b200.b200_ctx_create(C.byref(ctx), C.byref(g), 6, 2, -1))
    try:
        dpb = [synth.noise_planes(rng, W, H, bd) for _ in range(4)]
        for s in range(4): vvdec_b200.check(b200.b200_ctx_load_slot(ctx, s, abi.plane_ptrs(dpb[s])))
        for k, lm in enumerate((False, True)):
            pic = synth.gen_picture(rng, W, H, bd, dst_slot=4 + k, wp=True, lmcs=lm, pu_kw=dict(p_dmvr=0.0, p_bdof=0.0, p_bcw=0.3, p_geo=0.15))
            want, _ = oracle_decompress(oracle, g, dpb, pic)
            h = b200.b200_decompress_picture(ctx, C.byref(pic["struct"])); assert h >= 0, b200.b200_last_error()
            vvdec_b200.check(b200.b200_wait_picture(ctx, h, None, 0))
            got = [np.zeros_like(p) for p in want]
            vvdec_b200.check(b200.b200_get_frame(ctx, 4 + k, abi.plane_ptrs(got)))
            for c in range(3):
                assert np.array_equal(want[c], got[c]), f"picture {k} plane {c}: {len(np.argwhere(want[c] != got[c]))} diffs"
    finally:
        b200.b200_ctx_destroy(ctx)


@pytest.mark.parametrize("W,H,ctu,intra_frac,seed", [(416, 240, 128, 0.25, 1), (416, 240, 64, 1.0, 2), (832, 480, 128, 0.15, 3), (1920, 1080, 128, 0.3, 4)])
def test_picture_with_intra_cus(b200, oracle, W, H, ctu, intra_frac, seed):
    """Intra CUs reconstructed on the device inside the picture chain (regular modes): K2 for the inter CUs, K1 (inter TUs reconstruct, TUs
    of intra CUs leave their residual in the residual planes), K6 over the intra blocks in decoding order — each reads the reconstruction of inter and
    earlier intra neighbours — then deblocking / SAO / ALF.  intra_frac 1.0 is an I picture."""
    rng = np.random.default_rng(seed)
    bd = 10
    g = abi.make_geom(W, H, bd, ctu=ctu)
    ctx = C.c_void_p()
    vvdec_b200.check(b200.b200_ctx_create(C.byref(ctx), C.byref(g), 6, 2, -1))
    try:
        dpb = [synth.noise_planes(rng, W, H, bd) for _ in range(4)]
        for s in range(4): vvdec_b200.check(b200.b200_ctx_load_slot(ctx, s, abi.plane_ptrs(dpb[s])))
        for i in range(2):
            pic = synth.gen_picture(rng, W, H, bd, ctu=ctu, dst_slot=4 + i, intra_frac=intra_frac)
            assert len(pic["intraTus"]) > 0 and (pic["tus"]["flags"] & abi.TU_RESI).any() and (pic["intraTus"]["flags"] & abi.INTRA_ADD_RESI).any()
            assert intra_frac == 1.0 or (pic["intraTus"]["ciip"] > 0).any()          # CIIP CUs among the inter CUs
            want, dm_want = oracle_decompress(oracle, g, dpb, pic)
            h = b200.b200_decompress_picture(ctx, C.byref(pic["struct"])); assert h >= 0, b200.b200_last_error()
            dm = np.zeros((pic["ndmvr"] + 1, 2), np.int32)
            vvdec_b200.check(b200.b200_wait_picture(ctx, h, dm.ctypes.data, len(dm)))
            got = [np.zeros_like(p) for p in want]
            vvdec_b200.check(b200.b200_get_frame(ctx, 4 + i, abi.plane_ptrs(got)))
            for c in range(3):
                assert np.array_equal(want[c], got[c]), f"picture {i} plane {c}: {len(np.argwhere(want[c] != got[c]))} diffs"
            assert np.array_equal(dm, dm_want)
        # a record whose availability reaches outside the picture is refused
        bad = pic["intraTus"]; keep = bad[0].copy(); bad[0]["numAbove"] = 3; bad[0]["y"] = 0
        assert b200.b200_decompress_picture(ctx, C.byref(pic["struct"])) == -2 and b"intra block record" in b200.b200_last_error()
        bad[0] = keep
    finally:
        b200.b200_ctx_destroy(ctx)
