"""K6 intra prediction on the device (b200_intra_predict / b200_intra_reconstruct) against the oracle, which tests/test_intra_oracle_vs_ref.py
pins to the real IntraPrediction, and against the golden all-intra picture produced by the reference itself.  Whole pictures are predicted as one
list: every block reads what the blocks before it produced, so a single wrong sample (or a missed dependency) spreads over the picture."""
import ctypes as C
import os
import numpy as np
import pytest
import vvdec_b200
from vvdec_b200 import abi, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "k6_intra_picture.npz")


@pytest.mark.parametrize("W,H,bd,ctu,min_size,p_resi,seed", [(256, 128, 10, 128, 8, 0.0, 1), (192, 128, 10, 64, 4, 0.5, 2), (416, 240, 8, 128, 8, 0.5, 3),
                                                             (832, 480, 10, 128, 4, 0.3, 4), (1920, 1080, 10, 128, 8, 0.5, 5), (256, 256, 12, 64, 4, 1.0, 6)])
def test_intra_picture_vs_oracle(b200, oracle, W, H, bd, ctu, min_size, p_resi, seed):
    rng = np.random.default_rng(seed)
    g = abi.make_geom(W, H, bd, ctu=ctu)
    layout = synth.gen_intra_layout(rng, W, H, ctu, min_size=min_size)
    recs = synth.gen_intra_records(rng, layout, W, H, p_resi=p_resi, p_lm=0.25, colloc=seed & 1)
    planes = synth.noise_planes(rng, W, H, bd)
    resi = [rng.integers(-40, 41, size=p.shape).astype(np.int16) for p in planes]
    want = [p.copy() for p in planes]; got = [p.copy() for p in planes]
    oracle.orc_intra_reconstruct(C.byref(g), abi.plane_ptrs(want), abi.plane_ptrs(resi), recs.ctypes.data, len(recs))
    vvdec_b200.check(b200.b200_intra_reconstruct(C.byref(g), abi.plane_ptrs(got), abi.plane_ptrs(resi), recs.ctypes.data, len(recs)))
    for c in range(3):
        bad = np.argwhere(got[c] != want[c])
        assert len(bad) == 0, (c, len(bad), bad[:4].tolist())
    assert len(np.unique(recs["mode"])) > 40 and (recs["multiRefIdx"] > 0).any() and (recs["mode"] == 67).any() and (recs["mode"] == abi.INTRA_MIP).any()
    assert all((recs["mode"] == m).any() for m in (abi.INTRA_LM, abi.INTRA_MDLM_L, abi.INTRA_MDLM_T))


@pytest.mark.parametrize("W,H,bd,ctu,min_size,seed", [(256, 128, 10, 128, 4, 11), (416, 240, 8, 64, 4, 12), (832, 480, 10, 128, 8, 13), (1920, 1080, 10, 128, 4, 14)])
def test_intra_sub_partitions_vs_oracle(b200, oracle, W, H, bd, ctu, min_size, seed):
    """ISP CUs among regular ones (one record per prediction region, thin regions, per-TU residual masks): the small pictures run the one-CTA-per-block
    kernel, the dense 1080p list the CTU-resident one."""
    rng = np.random.default_rng(seed)
    g = abi.make_geom(W, H, bd, ctu=ctu)
    layout = synth.gen_intra_layout(rng, W, H, ctu, min_size=min_size)
    recs = synth.gen_intra_records(rng, layout, W, H, p_resi=0.5, p_lm=0.2, p_isp=0.4)
    isp = recs[(recs["flags"] & abi.INTRA_ISP) != 0]
    assert len(isp) > 20 and ((isp["mip"] & 3) == 2).any()
    if min_size == 4: assert (isp["log2h"] == 0).any() and (isp["ciip"] > 1).any() and (((isp["mip"] >> 4) & 3) == 0).any()      # 1-high regions, several TUs in a region, 4-wide CUs
    planes = synth.noise_planes(rng, W, H, bd)
    resi = [rng.integers(-40, 41, size=p.shape).astype(np.int16) for p in planes]
    want = [p.copy() for p in planes]; got = [p.copy() for p in planes]
    oracle.orc_intra_reconstruct(C.byref(g), abi.plane_ptrs(want), abi.plane_ptrs(resi), recs.ctypes.data, len(recs))
    vvdec_b200.check(b200.b200_intra_reconstruct(C.byref(g), abi.plane_ptrs(got), abi.plane_ptrs(resi), recs.ctypes.data, len(recs)))
    for c in range(3):
        bad = np.argwhere(got[c] != want[c])
        assert len(bad) == 0, (c, len(bad), bad[:4].tolist())
    # a region whose predecessor is missing is refused
    k = int(np.flatnonzero(((recs["flags"] & abi.INTRA_ISP) != 0) & (((recs["mip"] >> 2) & 3) == 1))[0])
    broken = np.delete(recs, k - 1)
    assert b200.b200_intra_reconstruct(C.byref(g), abi.plane_ptrs(got), abi.plane_ptrs(resi), broken.ctypes.data, len(broken)) == -2 and b"ISP" in b200.b200_last_error()


def test_intra_predict_only_and_golden(b200, oracle):
    z = np.load(GOLD)
    W, H, bd, ctu = [int(v) for v in z["geom"]]
    g = abi.make_geom(W, H, bd, ctu=ctu)
    src = [np.ascontiguousarray(z[f"src{c}"]) for c in range(3)]; resi = [np.ascontiguousarray(z[f"resi{c}"]) for c in range(3)]
    recs = np.ascontiguousarray(z["recs"])
    got = [p.copy() for p in src]
    vvdec_b200.check(b200.b200_intra_reconstruct(C.byref(g), abi.plane_ptrs(got), abi.plane_ptrs(resi), recs.ctypes.data, len(recs)))
    for c in range(3): assert np.array_equal(got[c], z[f"out{c}"]), c
    # prediction only (no residual planes): the flag is ignored
    want = [p.copy() for p in src]; got = [p.copy() for p in src]
    oracle.orc_intra_predict(C.byref(g), abi.plane_ptrs(want), recs.ctypes.data, len(recs))
    vvdec_b200.check(b200.b200_intra_predict(C.byref(g), abi.plane_ptrs(got), recs.ctypes.data, len(recs)))
    for c in range(3): assert np.array_equal(got[c], want[c]), c


def test_intra_argument_checks(b200):
    g = abi.make_geom(256, 128, 10)
    planes = [np.zeros((128, 256), np.int16), np.zeros((64, 128), np.int16), np.zeros((64, 128), np.int16)]
    r = np.zeros(1, abi.INTRA_TU_DTYPE)
    r["x"], r["y"], r["log2w"], r["log2h"], r["mode"] = 248, 0, 4, 4, 0
    assert b200.b200_intra_predict(C.byref(g), abi.plane_ptrs(planes), r.ctypes.data, 1) == -2 and b"geometry" in b200.b200_last_error()
    r["x"], r["numAbove"] = 0, 3
    assert b200.b200_intra_predict(C.byref(g), abi.plane_ptrs(planes), r.ctypes.data, 1) == -2 and b"availability" in b200.b200_last_error()
