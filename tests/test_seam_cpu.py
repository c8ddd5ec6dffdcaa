"""The DecLibRecon seam, executed on the CPU (SURVEY 8c level L1): a synthetic PARSED picture (oracle/ref_seam.h — real CodingStructure built through
the reference's Partitioner / addCU / addTU, levels in the reconstruction plane, motion as merge / AMVP syntax) is reconstructed by
  (i)  the reference's own DecLibRecon::decompressPicture / waitForPrevDecompressedPic (DecLibRecon.cpp:429,684) on a ThreadPool, and
  (ii) the host stages of the drop-in class b200glue::DecLibReconB200 (MIDER, boundary strengths, CU / TU walk -> work lists) followed by the oracle chain
       (K2 -> K1 -> K6 -> LMCS -> K3 -> K4 -> K5) on those lists — what the device executes (tests/test_seam_gpu.py runs the same pictures on the GPU).
Bit-exact equality of the three planes is required.  This pins the flatten walk over a parsed picture and the oracle chain at picture level against the
unmodified reference's picture-level entry point."""
import numpy as np, pytest
from tests import helpers

ref = helpers.load_ref()
pytestmark = pytest.mark.skipif(ref is None or not hasattr(ref, "ref_seam_create"), reason="oracle/_ref not built")

T_INTER = helpers.SEAM_INTER_TOOLS | helpers.SEAM_RESI_TOOLS | helpers.SEAM_FILTERS
ALL_TOOLS = helpers.SEAM_INTER_TOOLS | helpers.SEAM_RESI_TOOLS | helpers.SEAM_INTRA_TOOLS | helpers.SEAM_FILTERS
CASES = {
    "B_mixed_intra": dict(),                                                    # 15 % intra CUs, CIIP, GEO, affine, MMVD, SBT, MTS, LFNST, MIP, CCLM, full filter chain
    "B_intra_heavy": dict(intra=45, skip=5),
    "I_picture": dict(slice_type=2),
    "P_picture": dict(slice_type=1),
    "B_lmcs_inter": dict(lmcs=True, intra=0, tools=T_INTER),                    # LMCS with chroma scaling; inter CUs only
    "B_lmcs_intra_ciip": dict(lmcs=True),                                       # ... with intra and CIIP CUs (mapped-domain intra, luma-first ordering of the chroma scales)
    "I_lmcs": dict(lmcs=True, slice_type=2),
    "B_isp": dict(isp=40),                                                      # intra sub-partitions among the intra CUs
    "I_isp_lmcs": dict(isp=60, slice_type=2, lmcs=True),
    "B_3slices": dict(slices=3),                                                # per-slice reference lists, deblocking offsets, ALF APS lists / switches
    "B_4slices_lmcs_isp": dict(slices=4, lmcs=True, isp=30),
    "P_5slices": dict(slices=5, slice_type=1),
    "B_5slices_no_lf_across": dict(slices=5, ctu=64, lf_across_slices=False),    # ALF clipped sides / padded raster-slice corners, SAO / deblocking stop at slices
    "I_6slices_no_lf_across": dict(slices=6, ctu=64, lf_across_slices=False, slice_type=2),
    "B_weighted_prediction": dict(wp=True),                                     # explicit weights per slice: the glue's getWpScaling tables
    "P_wp_3slices": dict(wp=True, slice_type=1, slices=3),
    "B_local_dual_tree_isp": dict(tools=ALL_TOOLS | helpers.SEAM["LOCAL_DUAL_TREE"], isp=40),   # chroma-tree CUs, 4xN luma CUs (single-region ISP)
    "I_local_dual_tree_isp_lmcs": dict(tools=ALL_TOOLS | helpers.SEAM["LOCAL_DUAL_TREE"], isp=40, slice_type=2, lmcs=True),
    "B_scaling_lists": dict(scaling_lists=True),                                # explicit scaling lists: the reference's table set as the picture's scaling arena
    "I_scaling_lists_isp_lmcs": dict(scaling_lists=True, slice_type=2, isp=30, lmcs=True),
    "B_ctu64": dict(ctu=64),
    "B_ctu32_8bit": dict(ctu=32, bd=8),
    "B_no_dmvr": dict(tools=(helpers.SEAM_INTER_TOOLS | helpers.SEAM_RESI_TOOLS | helpers.SEAM_INTRA_TOOLS | helpers.SEAM_FILTERS) & ~helpers.SEAM["DMVR"]),
    "B_no_filters": dict(deblock=False, tools=helpers.SEAM_INTER_TOOLS | helpers.SEAM_RESI_TOOLS | helpers.SEAM_INTRA_TOOLS),
}


def run_case(seed, W, H, threads, **kw):
    oracle = helpers.load_oracle()
    case = helpers.SeamCase(ref, np.random.default_rng(seed), W, H, **kw)
    out, col, secs = case.run_stock(threads=threads)
    pic, s2 = case.flatten(threads=threads)
    assert pic is not None, f"DecLibReconB200 refused the picture ({s2})"
    want, _ = helpers.oracle_decompress(oracle, case.g, case.refs + [[np.zeros_like(p) for p in case.refs[0]]], pic)
    for c in range(3):
        assert np.array_equal(want[c], out[c]), f"plane {c}: {np.count_nonzero(want[c] != out[c])} samples differ from the reference's DecLibRecon"
    if not (case.cfg.tools & helpers.SEAM["DMVR"]) or case.cfg.sliceType != 0:
        # collocated motion (what TaskFinishMotionInfo leaves for later pictures' TMVP); a dry run has no DMVR deltas, so only pictures without DMVR compare here
        assert helpers.col_motion_diff(col, pic["colMotion"], case.g) == 0
    return case


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_stock_declibrecon_vs_flatten_oracle(name, seed):
    run_case(seed, 416, 240, 0, **CASES[name])


def test_tool_coverage_of_the_generator():
    st = helpers.SeamCase(ref, np.random.default_rng(7), 832, 480).stats()
    for k in ("intra", "skip", "merge", "affine", "geo", "ciip", "mmvd", "resi", "sbt", "lfnst", "mts", "mip"):
        assert st[k] > 0, (k, st)
    assert 0.05 < st["intra"] / st["cus"] < 0.3, st


def test_thread_pool_runs_match_single_thread():
    """Both back ends on a pool of 4 threads (wave-front MIDER rows, parallel flatten rows) give what the single-threaded runs give."""
    run_case(11, 832, 480, 4)
    run_case(12, 832, 480, 4, slice_type=2)


def test_1080p_picture():
    run_case(21, 1920, 1080, 4)


def test_unsupported_tool_follows_the_error_contract():
    """A picture whose header announces virtual boundaries: the stock back end reconstructs it; DecLibReconB200 throws UnsupportedFeatureException,
    which must surface as pic->error + reconDone exception (DecLibRecon.cpp:704-715) — rc -4 here — and leave the recon object usable for the next picture."""
    case = helpers.SeamCase(ref, np.random.default_rng(5), 416, 240, virtual_boundaries=True)
    case.run_stock(threads=0)
    pic, rc = case.flatten(threads=0)
    assert pic is None and rc == -4.0
    run_case(6, 416, 240, 0)            # the same (static) recon object afterwards


def test_two_alternating_recon_instances():
    """DecLib keeps two recon instances on one thread pool and hands them pictures in turn (DecLib.h:70): the stock back end run that way gives what the
    one-picture-at-a-time runs give, and two DecLibReconB200 instances (dry run: host stages, shared DPB bookkeeping) get through the same pictures."""
    rng = np.random.default_rng(77)
    base = helpers.SeamCase(ref, rng, 416, 240)
    cases = [base.variant(seed=100 + i, slice_type=2 if i == 2 else 0) for i in range(5)]
    secs, outs = helpers.seam_pipelined(ref, cases, 4, 0, 2)
    assert secs >= 0 and len(outs) == 5
    for c, (out, col) in zip(cases, outs):
        want, col_want, _ = c.run_stock(threads=4)
        for k in range(3): assert np.array_equal(want[k], out[k])
        assert helpers.col_motion_diff(col_want, col, c.g) == 0
    secs, _ = helpers.seam_pipelined(ref, cases, 4, 2, 2)
    assert secs >= 0


@pytest.mark.parametrize("threads", [0, 1, 4])
def test_pictures_completed_by_a_pool_task(threads):
    """setAsyncFinish: the class releases pic->reconDone from a task of its own, as the reference's finishReconTask does, instead of inside
    waitForPrevDecompressedPic() — two alternating instances, chained pictures (dry run), with no, one and four pool threads."""
    rng = np.random.default_rng(78)
    base = helpers.SeamCase(ref, rng, 416, 240)
    cases = [base.variant(seed=200 + i, slice_type=2 if i == 0 else 0) for i in range(6)]
    ref.ref_seam_set_async_finish(1)
    try:
        secs, _ = helpers.seam_pipelined(ref, cases, threads, 2, 2, read=False, chain=True)
    finally:
        ref.ref_seam_set_async_finish(0)
    assert secs >= 0


def test_reference_still_in_flight_on_the_other_instance():
    """Picture B predicts from picture A while A is still with the other recon instance (what DecLib's two alternating instances produce all the time).  Stock: the two
    instances give B the same samples as one instance after the other.  DecLibReconB200 (dry run): B's submission waits until A is in the stream, B's lists name A's
    DPB slot, and the oracle chain A -> B over those lists reproduces the stock pictures."""
    oracle = helpers.load_oracle()
    rng = np.random.default_rng(91)
    base = helpers.SeamCase(ref, rng, 416, 240, intra=10)
    A, B = base.variant(seed=301), base.variant(seed=302)
    secs, outs = helpers.seam_pipelined(ref, [A, B], 4, 0, 2, chain=True)
    assert secs >= 0
    secs, _, lists = helpers.seam_pipelined(ref, [A, B], 4, 2, 2, read=False, chain=True, flat=True)
    assert secs >= 0
    fa, fb = lists
    slotA, slotB = fa["struct"].dstSlot, fb["struct"].dstSlot
    assert slotA != slotB and (fb["pus"]["refSlot"] == slotA).any(), "B's lists do not reference A's slot"
    zeros = [np.zeros_like(p) for p in base.refs[0]]
    dpb = [base.refs[0], base.refs[1], base.refs[2], base.refs[3]] + [zeros] * 4
    wantA, _ = helpers.oracle_decompress(oracle, base.g, dpb, fa)
    for c in range(3): assert np.array_equal(wantA[c], outs[0][0][c]), f"A plane {c}"
    dpb[slotA] = wantA
    wantB, _ = helpers.oracle_decompress(oracle, base.g, dpb, fb)
    for c in range(3): assert np.array_equal(wantB[c], outs[1][0][c]), f"B plane {c}: {np.count_nonzero(wantB[c] != outs[1][0][c])} samples differ"

