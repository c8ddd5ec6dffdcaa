"""LMCS: the oracle's restatement (oracle/k6_lmcs.c) and the generator's table construction against the reference's real Reshape class
and PelBufferOps pointers (scalar and SIMD), function level and picture level."""
import ctypes as C
import numpy as np
import pytest
from vvdec_b200 import abi, synth
from tests.helpers import aligned, aligned_copy, ref_ptrs, oracle_decompress

pytestmark = pytest.mark.ref


def build(ref, rng, bd, chroma_adj=True):
    """random legal model -> (Lmcs from the REAL constructReshaper, generator's dict)"""
    cus = np.array([[0, 0, 64, 64]])
    m = synth.gen_lmcs(rng, bd, cus, 64, 64, 64, chroma_adj=chroma_adj)
    L = abi.Lmcs(); lut = np.zeros(1 << bd, np.int16)
    delta = (C.c_int * 16)(*m["delta"])
    assert ref.ref_lmcs_build(bd, m["minBin"], m["maxBin"], delta, m["chrOff"], int(chroma_adj), C.byref(L), lut) == 0
    return L, lut, m


@pytest.mark.parametrize("bd", [8, 10, 12])
def test_tables_match_construct_reshaper(ref, bd):
    rng = np.random.default_rng(bd)
    for _ in range(20):
        L, lut, m = build(ref, rng, bd)
        G = m["struct"]
        assert (L.orgCW, L.minBinIdx, L.maxBinIdx) == (G.orgCW, G.minBinIdx, G.maxBinIdx)
        assert list(L.reshapePivot) == list(G.reshapePivot) and list(L.inputPivot) == list(G.inputPivot)
        assert list(L.fwdScaleCoef) == list(G.fwdScaleCoef) and list(L.chromaAdjHelpLUT) == list(G.chromaAdjHelpLUT)
        assert np.array_equal(lut, m["invLUT"])


@pytest.mark.parametrize("simd", [0, 1])
@pytest.mark.parametrize("bd", [8, 10, 12])
def test_forward_inverse_scale(oracle, ref, bd, simd):
    rng = np.random.default_rng(100 + bd + simd)
    for (w, h) in [(4, 4), (8, 4), (16, 16), (64, 32), (128, 128), (4, 64)]:
        L, lut, m = build(ref, rng, bd)
        st = (w + 15) // 16 * 16 + 16                                # the AVX2 paths use aligned 32-byte row loads (picture buffers are)
        src = rng.integers(0, 1 << bd, size=(h, st)).astype(np.int16); src[0, :4] = [0, (1 << bd) - 1, 1, (1 << bd) - 2]
        a = aligned_copy(src); b = aligned_copy(src)
        ref.ref_lmcs_fwd_block(simd, a.ctypes.data, st, w, h)
        oracle.orc_lmcs_fwd_block(b.ctypes.data, st, w, h, bd, C.byref(L))
        assert np.array_equal(a, b), ("fwd", w, h)
        a = aligned_copy(src); b = aligned_copy(src)
        ref.ref_lmcs_inv_block(simd, a.ctypes.data, st, w, h)          # SIMD: the piece-wise linear rspBcw; scalar: applyLut(m_invLUT)
        b[:, :w] = lut[b[:, :w]]
        assert np.array_equal(a, b), ("inv", w, h)
        # scaleSignal: residuals incl. the extremes, every LUT entry
        res = rng.integers(-(1 << bd) - 40, (1 << bd) + 40, size=(h, w)).astype(np.int16); res[0, :2] = [-32768, 32767]
        for sc in sorted(set(L.chromaAdjHelpLUT)):
            a = res.copy()
            ref.ref_lmcs_scale_block(a.ctypes.data, w, w, h, sc, bd)
            want = np.array([oracle.orc_lmcs_scale_resi(int(v), sc, bd) for v in res.reshape(-1)], np.int16).reshape(h, w)
            assert np.array_equal(a, want), ("scale", sc)


@pytest.mark.parametrize("ctu,W,H", [(128, 256, 192), (64, 192, 136), (32, 96, 72)])
def test_vpdu_chroma_scale(oracle, ref, ctu, W, H):
    """calculateChromaAdjVpduNei on a picture with one CU per CTU: every VPDU, incl. the picture-edge clamps of the neighbour walk."""
    bd = 10
    rng = np.random.default_rng(ctu)
    g = abi.make_geom(W, H, bd, ctu=ctu)
    L, lut, m = build(ref, rng, bd)
    planes = synth.noise_planes(rng, W, H, bd)
    vs = 64 if ctu == 128 else ctu
    for vy in range(0, H, vs):
        for vx in range(0, W, vs):
            cx, cy = vx // ctu * ctu, vy // ctu * ctu          # the CU covering the VPDU's top-left = its CTU
            v = abi.LmcsVpdu(cx, cy, int(cx > 0), int(cy > 0))
            want = ref.ref_lmcs_vpdu_scale(C.byref(g), abi.plane_ptrs(planes), vx, vy)
            got = oracle.orc_lmcs_vpdu_scale(C.byref(g), planes[0], C.byref(L), C.byref(v))
            assert got == want, (vx, vy)


@pytest.mark.parametrize("simd", [0, 1])
@pytest.mark.parametrize("chroma_adj", [True, False])
def test_picture_with_lmcs(oracle, ref, simd, chroma_adj):
    """Whole back end with LMCS on: oracle chain vs the reference's kernels (real rspBufFwd / calculateChromaAdjVpduNei / scaleSignal /
    rspBcw|applyLut inside the multi-threaded reference arm).  The reference arm's structure has one CU per CTU, so the VPDU records
    point at CTU origins here."""
    W, H, bd, ctu = 256, 192, 10, 128
    rng = np.random.default_rng(7 + simd)
    g = abi.make_geom(W, H, bd, ctu=ctu)
    dpb = [synth.noise_planes(rng, W, H, bd) for _ in range(4)]
    pic = synth.gen_picture(rng, W, H, bd, dst_slot=0, lmcs=True, lmcs_chroma=chroma_adj)
    vp = pic["lmcs"]["vpdus"]
    for j in range((H + 63) // 64):
        for i in range((W + 63) // 64):
            cx, cy = i * 64 // ctu * ctu, j * 64 // ctu * ctu
            vp[j * ((W + 63) // 64) + i] = (cx, cy, cx > 0, cy > 0)
    want, _ = oracle_decompress(oracle, g, dpb, pic)
    got = [np.zeros_like(p) for p in want]
    ref.ref_decompress_picture_out(C.byref(g), ref_ptrs(dpb), C.byref(pic["struct"]), 3, simd, abi.plane_ptrs(got))
    for c in range(3):
        assert np.array_equal(want[c], got[c]), f"plane {c}: {len(np.argwhere(want[c] != got[c]))} diffs"
