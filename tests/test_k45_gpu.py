"""K4 (SAO) and K5 (ALF + CC-ALF) parity on the GPU: CUDA kernels (through the C ABI) vs the pinned oracle, bit-exact."""
import ctypes as C
import numpy as np
import pytest
import vvdec_b200
from vvdec_b200 import abi, synth

pytestmark = pytest.mark.gpu
CASES = [(1, 256, 128, 10, 128), (2, 416, 240, 10, 64), (3, 200, 136, 8, 32), (4, 1920, 1080, 10, 128), (5, 384, 256, 12, 128),
         (6, 3840, 2160, 10, 128)]


@pytest.mark.parametrize("seed,W,H,bd,ctu", CASES)
@pytest.mark.parametrize("vb", [0, 1])
def test_sao_gpu_vs_oracle(b200, oracle, seed, W, H, bd, ctu, vb):
    rng = np.random.default_rng(seed)
    src = synth.noise_planes(rng, W, H, bd)
    sao = synth.gen_sao(rng, W, H, ctu, bd, p_on=0.7)
    sao["avail"] = np.where(rng.random(len(sao)) < 0.3, sao["avail"] & rng.integers(0, 256, size=len(sao)).astype(np.uint8), sao["avail"])
    g = abi.make_geom(W, H, bd, ctu=ctu)
    v = abi.Vb()
    if vb:
        v.numVer, v.numHor = 2, 1
        v.posX[0], v.posX[1], v.posY[0] = 8 * (W // 24), 8 * (W // 12), 8 * (H // 16)
    a = [np.zeros_like(p) for p in src]; b = [np.zeros_like(p) for p in src]
    oracle.orc_sao_picture(C.byref(g), abi.plane_ptrs(src), abi.plane_ptrs(a), sao.ctypes.data, C.addressof(v))
    vvdec_b200.check(b200.b200_sao_picture(C.byref(g), abi.plane_ptrs(src), abi.plane_ptrs(b), sao.ctypes.data, C.addressof(v)))
    for c in range(3):
        assert np.array_equal(a[c], b[c]), f"plane {c}: {np.argwhere(a[c] != b[c])[:8]}"
        assert not np.array_equal(a[c], src[c])


@pytest.mark.parametrize("seed,W,H,bd,ctu", [c for c in CASES if c[3] <= 10])
def test_alf_gpu_vs_oracle(b200, oracle, seed, W, H, bd, ctu):
    rng = np.random.default_rng(seed)
    src = synth.noise_planes(rng, W, H, bd)
    t = synth.gen_alf(rng, W, H, ctu, bd, n_aps=3)
    T = abi.make_alf_tables(t)
    g = abi.make_geom(W, H, bd, ctu=ctu)
    a = [np.zeros_like(p) for p in src]; b = [np.zeros_like(p) for p in src]
    oracle.orc_alf_picture(C.byref(g), abi.plane_ptrs(src), abi.plane_ptrs(a), t["ctus"].ctypes.data, C.byref(T))
    vvdec_b200.check(b200.b200_alf_picture(C.byref(g), abi.plane_ptrs(src), abi.plane_ptrs(b), t["ctus"].ctypes.data, C.byref(T)))
    for c in range(3):
        assert np.array_equal(a[c], b[c]), f"plane {c}: {len(np.argwhere(a[c] != b[c]))} diffs, first {np.argwhere(a[c] != b[c])[:8]}"
        assert not np.array_equal(a[c], src[c])


@pytest.mark.parametrize("seed,W,H,bd,ctu", [(41, 416, 240, 10, 64), (42, 832, 480, 10, 128), (43, 256, 256, 8, 32)])
def test_alf_clipped_ctus_and_padded_corners(b200, oracle, seed, W, H, bd, ctu):
    """CTUs whose neighbours ALF may not read (in-loop filtering disabled across slices / tiles, AdaptiveLoopFilter.cpp:763-848): every combination of clipped sides,
    raster-slice corner padding in both forms (chroma padded with its own or with the luma margin).  The oracle's rule is pinned against the stock decoder through
    the seam (tests/test_seam_cpu.py, slices with loop filtering across them disabled)."""
    rng = np.random.default_rng(seed)
    src = synth.noise_planes(rng, W, H, bd)
    t = synth.gen_alf(rng, W, H, ctu, bd, n_aps=3, p_luma=0.9, p_chroma=0.8)
    n = len(t["ctus"])
    f = rng.integers(0, 16, size=n).astype(np.uint8) << 1                               # CLIP_TOP / BOTTOM / LEFT / RIGHT
    f = np.where(rng.random(n) < 0.3, 0, f)
    ptl = (rng.random(n) < 0.5) & ((f & (abi.ALF_CLIP_TOP | abi.ALF_CLIP_LEFT)) == 0)
    pbr = (rng.random(n) < 0.5) & ((f & (abi.ALF_CLIP_BOTTOM | abi.ALF_CLIP_RIGHT)) == 0)
    # the reference pads a corner only where that diagonal CTU exists
    ctusW = (W + ctu - 1) // ctu; ctusH = (H + ctu - 1) // ctu; idx = np.arange(n)
    ptl &= (idx % ctusW > 0) & (idx // ctusW > 0); pbr &= (idx % ctusW < ctusW - 1) & (idx // ctusW < ctusH - 1)
    f = f | (ptl * abi.ALF_PAD_TL).astype(np.uint8) | (pbr * abi.ALF_PAD_BR).astype(np.uint8)
    t["ctus"]["enable"][:, 0] |= f
    wide = (rng.random((n, 2)) < 0.5) & (f != 0)[:, None]
    t["ctus"]["ccIdx"][wide] = 0                                                        # the wide form belongs to slices without CC-ALF for that component
    t["ctus"]["enable"][:, 1:][wide] |= abi.ALF_PAD_WIDE
    assert ptl.any() and pbr.any() and (f & 30).any() and wide.any()
    T = abi.make_alf_tables(t)
    g = abi.make_geom(W, H, bd, ctu=ctu)
    a = [np.zeros_like(p) for p in src]; b = [np.zeros_like(p) for p in src]
    oracle.orc_alf_picture(C.byref(g), abi.plane_ptrs(src), abi.plane_ptrs(a), t["ctus"].ctypes.data, C.byref(T))
    vvdec_b200.check(b200.b200_alf_picture(C.byref(g), abi.plane_ptrs(src), abi.plane_ptrs(b), t["ctus"].ctypes.data, C.byref(T)))
    for c in range(3):
        assert np.array_equal(a[c], b[c]), f"plane {c}: {len(np.argwhere(a[c] != b[c]))} diffs, first {np.argwhere(a[c] != b[c])[:8]}"
