"""The DecLibRecon seam on the GPU (SURVEY 8c level L1): the reference's own DecLibRecon (CPU, ThreadPool) and the drop-in class
b200glue::DecLibReconB200 (device, through the C ABI) reconstruct the SAME synthetic parsed Picture — same create / decompressPicture /
waitForPrevDecompressedPic calls (oracle/ref_seam.h).  Required bit-exact: the three reconstruction planes as they land in Picture::m_bufs and the
collocated motion field (CodingStructure::m_colMiMap: what TaskFinishMotionInfo leaves for later pictures' TMVP, DMVR refinements included)."""
import numpy as np, pytest
from tests import helpers

ref = helpers.load_ref()
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(ref is None or not hasattr(ref, "ref_seam_create"), reason="oracle/_ref not built")]

T_INTER = helpers.SEAM_INTER_TOOLS | helpers.SEAM_RESI_TOOLS | helpers.SEAM_FILTERS


def both(seed, W, H, threads=4, **kw):
    case = helpers.SeamCase(ref, np.random.default_rng(seed), W, H, **kw)
    want, col_want, t_cpu = case.run_stock(threads=threads)
    got, col_got, t_gpu = case.run_b200(threads=threads)
    assert t_gpu >= 0, f"DecLibReconB200 failed ({t_gpu})"
    for c in range(3):
        assert np.array_equal(want[c], got[c]), f"plane {c}: {np.count_nonzero(want[c] != got[c])} samples differ"
    assert helpers.col_motion_diff(col_want, col_got, case.g) == 0, "collocated motion (colMotion / DMVR write-back) differs"
    return t_cpu, t_gpu


@pytest.mark.parametrize("name,kw", [("B_mixed_intra", dict()), ("I_picture", dict(slice_type=2)), ("P_picture", dict(slice_type=1)),
                                     ("B_lmcs_inter", dict(lmcs=True, intra=0, tools=T_INTER)), ("B_lmcs_intra_ciip", dict(lmcs=True)), ("I_lmcs", dict(lmcs=True, slice_type=2)),
                                     ("B_ctu64", dict(ctu=64)),
                                     ("B_3slices", dict(slices=3)), ("B_4slices_lmcs_isp", dict(slices=4, lmcs=True, isp=30)), ("I_2slices", dict(slices=2, slice_type=2)),
                                     ("B_5slices_no_lf_across", dict(slices=5, ctu=64, lf_across_slices=False)), ("B_7slices_no_lf_across_ctu32", dict(slices=7, ctu=32, lf_across_slices=False)),
                                     ("B_weighted_prediction", dict(wp=True)), ("P_wp_3slices", dict(wp=True, slice_type=1, slices=3)),
                                     ("B_local_dual_tree_isp", dict(tools=T_INTER | helpers.SEAM_INTRA_TOOLS | helpers.SEAM["LOCAL_DUAL_TREE"], isp=40)),
                                     ("I_local_dual_tree_isp", dict(tools=T_INTER | helpers.SEAM_INTRA_TOOLS | helpers.SEAM["LOCAL_DUAL_TREE"], isp=40, slice_type=2)),
                                     ("B_scaling_lists", dict(scaling_lists=True)), ("I_scaling_lists_isp", dict(scaling_lists=True, slice_type=2, isp=30)),
                                     ("B_isp", dict(isp=40)), ("I_isp", dict(isp=60, slice_type=2)), ("I_isp_lmcs", dict(isp=60, slice_type=2, lmcs=True)), ("I_isp_ctu32", dict(isp=70, slice_type=2, ctu=32))])
@pytest.mark.parametrize("seed", [1, 2])
def test_seam_small(name, kw, seed):
    both(seed, 416, 240, **kw)


def test_seam_1080p_and_4k():
    both(31, 1920, 1080, threads=8)
    both(32, 3840, 2160, threads=16)
    both(33, 3840, 2160, threads=16, lmcs=True)
    both(34, 3840, 2160, threads=16, lmcs=True, slice_type=2)
    both(35, 3840, 2160, threads=16, isp=50, slice_type=2)                   # dense list: the CTU-resident K6
    both(36, 3840, 2160, threads=16, isp=50)                                 # sparse list: one CTA per block
    both(37, 1920, 1080, threads=8, slices=6, isp=20)


def test_seam_error_contract_on_the_device_path():
    case = helpers.SeamCase(ref, np.random.default_rng(5), 416, 240, virtual_boundaries=True)
    _, _, rc = case.run_b200(threads=2)
    assert rc == -4.0
    both(6, 416, 240)


def test_two_alternating_recon_instances_on_the_device():
    """Two DecLibReconB200 instances on one pool and one device context take pictures in turn, as DecLib runs its recon instances (DecLib.h:70):
    every picture comes out as the stock back end reconstructs it."""
    rng = np.random.default_rng(78)
    base = helpers.SeamCase(ref, rng, 832, 480, isp=20)
    cases = [base.variant(seed=200 + i, slice_type=2 if i == 3 else 0) for i in range(7)]
    secs, outs = helpers.seam_pipelined(ref, cases, 4, 1, 2)
    assert secs >= 0 and len(outs) == 7
    for c, (out, col) in zip(cases, outs):
        want, col_want, _ = c.run_stock(threads=4)
        for k in range(3): assert np.array_equal(want[k], out[k]), f"plane {k}"
        assert helpers.col_motion_diff(col_want, col, c.g) == 0

