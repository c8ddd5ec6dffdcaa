"""The DecLibRecon seam from a real BITSTREAM, on the CPU (SURVEY 8c level L2, row f-4).

oracle/vvc_stream.py writes a VVC stream: headers field by field, slice data drawn by the reference's own CABACReader over a recording bin source and
arithmetic-encoded with the reference's probability models.  Every stream is decoded
  (i)   by the build that drew it (its reconstruction is what the stream means),
  (ii)  by the stock reference through its public API (vvdec_decode / vvdec_flush) — (i) == (ii) pins the writer / arithmetic encoder, and
  (iii) by the reference with b200glue::DecLibReconB200 compiled in behind the seam (oracle/_ref/libvvdec_swapped.so): parser, DecLib scheduling, picture
        recycling and output are the reference's, the drop-in class runs its host stages, and the oracle chain stands where the device would be
        (tests/test_stream_gpu.py runs the same streams on the GPU).
Bit-exact equality of all output frames is required."""
import os, numpy as np, pytest
from tests import helpers
from oracle import vvc_stream as vs

pytestmark = pytest.mark.skipif(not (vs.available() and os.path.exists(vs.SWAP_SO)), reason="oracle/_ref not built")

INTRA = dict(isp=True, mrl=True, mip=True, cclm=True, lfnst=True, mts=True, mts_intra=True, mts_inter=True, jccr=True, dep_quant=True, sign_hiding=True, sao=True)
INTER = dict(temporal_mvp=True, sbtmvp=True, amvr=True, bdof=True, smvd=True, dmvr=True, mmvd=True, sbt=True, affine=True, affine_6param=True, affine_amvr=True,
             prof=True, bcw=True, ciip=True, gpm=True)
ALL = {**INTRA, **INTER}


def gop4(base=0, idr=True):
    """random-access GOP of 4 in decoding order: I/P anchor, B at the middle, two non-reference Bs"""
    b = base
    first = vs.Pic(b, idr=True) if idr else vs.Pic(b, vs.SLICE_P, ((b - 4,), ()))
    return [first, vs.Pic(b + 4, vs.SLICE_P, ((b,), ())), vs.Pic(b + 2, vs.SLICE_B, ((b,), (b + 4,))),
            vs.Pic(b + 1, vs.SLICE_B, ((b, b + 2), (b + 2, b + 4)), referenced=False), vs.Pic(b + 3, vs.SLICE_B, ((b + 2, b), (b + 4,)), referenced=False)]


def gop8(base=0, idr=True, n_gops=1):
    """hierarchical-B random-access GOPs of 8 (decoding order 8 4 2 1 3 6 5 7 after the anchor), up to three references per list, non-reference pictures at the top level"""
    pics = []
    for g in range(n_gops):
        b = base + 8 * g
        if g == 0: pics.append(vs.Pic(b, idr=True) if idr else vs.Pic(b, vs.SLICE_P, ((b - 8,), ())))
        B = vs.SLICE_B
        pics += [vs.Pic(b + 8, vs.SLICE_P if g % 2 == 0 else B, ((b,), ()) if g % 2 == 0 else ((b,), (b,))),
                 vs.Pic(b + 4, B, ((b, b + 8), (b + 8, b))),
                 vs.Pic(b + 2, B, ((b, b + 4), (b + 4, b + 8))),
                 vs.Pic(b + 1, B, ((b, b + 2), (b + 2, b + 4, b + 8)), referenced=False),
                 vs.Pic(b + 3, B, ((b + 2, b), (b + 4, b + 8)), referenced=False),
                 vs.Pic(b + 6, B, ((b + 4, b + 2, b), (b + 8,))),
                 vs.Pic(b + 5, B, ((b + 4, b), (b + 6, b + 8)), referenced=False),
                 vs.Pic(b + 7, B, ((b + 6, b + 4), (b + 8,)), referenced=False)]
    return pics


def low_delay(n):
    """I P P P ... each picture predicting from the two before it; the last list-1 entry repeats list 0 (low-delay B)"""
    pics = [vs.Pic(0)]
    for p in range(1, n):
        refs = tuple(r for r in (p - 1, p - 2) if r >= 0)
        pics.append(vs.Pic(p, vs.SLICE_B if p % 3 else vs.SLICE_P, (refs, refs if p % 3 else ())))
    return pics


def _mixed_slice_types():
    pics = gop4()
    for q in pics[1:]: q["slice_types"] = [q.slice_type, vs.SLICE_I, q.slice_type]        # an intra slice between two inter slices
    return pics


def _weighted(pics):
    for i, q in enumerate(pics): q["wp"] = 40 + i
    return pics


def _picture_switches():
    pics = gop4()
    for i, q in enumerate(pics): q["bdof"], q["dmvr"], q["prof"], q["jccr_sign"] = bool(i & 1), bool(i & 2), bool((i + 1) & 2), bool(i & 1)
    return pics


SL3 = dict(width=256, height=256, slice_rows=(2, 1, 1))
def _intra_pictures_inside():
    P, B, I = vs.SLICE_P, vs.SLICE_B, vs.SLICE_I                      # non-IDR intra pictures (their reference lists are written, one of them empty)
    return [vs.Pic(0), vs.Pic(1, P, ((0,), ())), vs.Pic(2, B, ((1, 0), (1,))), vs.Pic(3, I, ((2, 1), (2,)), idr=False), vs.Pic(4, P, ((3, 2), ())),
            vs.Pic(5, I, ((), ()), idr=False), vs.Pic(6, B, ((5,), (5,)))]


CASES = {
    "I_all_intra_tools": (dict(INTRA), lambda: [vs.Pic(0)]),
    "I_dual_tree_ctu128": (dict(INTRA, ctu=128, dual_tree=True), lambda: [vs.Pic(0), vs.Pic(1, idr=True)]),
    "I_transform_skip_bdpcm": (dict(INTRA, transform_skip=True, bdpcm=True), lambda: [vs.Pic(0, dep_quant=False, sign_hiding=False)]),
    "I_sign_hiding": (dict(INTRA), lambda: [vs.Pic(0, dep_quant=False)]),
    "gop_no_inter_tools": (dict(INTRA), gop4),
    "gop_all_tools": (dict(ALL), gop4),
    "gop_all_tools_ctu128": (dict(ALL, ctu=128), gop4),
    "gop_ctu32_8bit": (dict(ALL, ctu=32, max_bt_inter=32, max_tt_inter=32, bit_depth=8), gop4),
    "gop_picture_not_ctu_aligned": (dict(ALL, width=200, height=104), gop4),
    "gop_cu_qp_delta": (dict(ALL, cu_qp_delta=True), gop4),
    "gop_min_cb8_qp20": (dict(ALL, min_cb=8, min_qt_intra=16, min_qt_inter=16, min_qt_intra_c=16, init_qp=20), gop4),
    "gop_no_deblocking": (dict(ALL, deblocking_disabled=True), gop4),
    "low_delay_8": (dict(ALL), lambda: low_delay(8)),
    "low_delay_with_intra_pictures": (dict(ALL), _intra_pictures_inside),
    "gop8_x2": (dict(ALL, dpb_size=8), lambda: gop8(n_gops=2)),
    "gop8_alf_lmcs_3slices": (dict(ALL, dpb_size=8, alf=True, ccalf=True, lmcs=True, **SL3), lambda: vs.with_lmcs(vs.with_alf(gop8(), np.random.default_rng(41)), np.random.default_rng(42), every=3)),
    "gop_max_transform_32": (dict(ALL, max_tb64=False), gop4),                                                      # 64x64 CUs carry four TUs; CIIP still predicts the CU block
    "gop_4tiles_one_slice": (dict(ALL, width=256, height=192, tiles=((2, 2), (1, 2))), gop4),                       # CABAC restarts per tile inside the slice data
    "gop_4tiles_4slices": (dict(ALL, width=256, height=192, tiles=((1, 3), (2, 1)), slice_per_tile=True), gop4),
    "gop_4tiles_no_lf_across_alf": (dict(ALL, width=256, height=192, tiles=((2, 2), (1, 2)), lf_across_tiles=False, alf=True, ccalf=True),
                                    lambda: vs.with_alf(gop4(), np.random.default_rng(21))),
    "gop_picture_level_tool_switches": (dict(ALL, ph_tool_control=True), _picture_switches),                        # ph_bdof / dmvr / prof _disabled_flag, ph_joint_cbcr_sign_flag
    "gop_chroma_qp_tables": (dict(ALL, chroma_qp_tables=((-4, [(1, 1), (3, 2), (0, 0)]), (2, [(2, 3)]), (-10, [(5, 2), (1, 2)])), chroma_qp_offsets=(1, -2, 3), slice_chroma_qp_offsets=True), gop4),
    "gop_ladf_chroma_deblock_offsets": (dict(ALL, ladf=(-3, [(2, 100), (-4, 200), (5, 150)]), chroma_qp_offsets=(0, 0, 0), cb_cr_deblock_offsets=(3, -2, -4, 5), beta_offset_div2=-2, tc_offset_div2=4), gop4),
    "gop_ts_min_qp_lfnst_without_scaling": (dict(ALL, transform_skip=True, bdpcm=True, min_qp_prime_ts=3, ts_max_size=4, scaling_lists=True, lfnst_scaling_disabled=True, parallel_merge_level=5),
                                            lambda: vs.with_scaling_lists(gop4(), np.random.default_rng(22))),
    "gop_3slices": (dict(ALL, **SL3), gop4),                                                                       # per-slice QP, SAO switches, reference order, dep. quant
    "gop_4slices_no_lf_across_deblock_override": (dict(ALL, width=256, height=256, slice_rows=(1, 1, 1, 1), lf_across_slices=False, deblocking_override=True), gop4),
    "gop_intra_slice_in_inter_pictures": (dict(ALL, **SL3), _mixed_slice_types),
    "gop_3slices_alf_lmcs": (dict(ALL, **SL3, alf=True, ccalf=True, lmcs=True, lf_across_slices=False),
                             lambda: vs.with_lmcs(vs.with_alf(gop4(), np.random.default_rng(11)), np.random.default_rng(12))),
    "gop_weighted_prediction": (dict(ALL, weighted_pred=True, weighted_bipred=True), lambda: _weighted(gop4())),
    "low_delay_weighted_3slices": (dict(ALL, weighted_pred=True, weighted_bipred=True, **SL3), lambda: _weighted(low_delay(5))),
    "gop_chroma_qp_offsets": (dict(ALL, chroma_qp_offsets=(2, -3, 1), slice_chroma_qp_offsets=True), gop4),
    "gop_cu_chroma_qp_offsets": (dict(ALL, chroma_qp_offsets=(1, -1, 0), cu_chroma_qp_offset_list=((2, -2, 1), (-3, 3, -1), (5, 4, 3)), cu_qp_delta=True), gop4),
    "gop_scaling_lists": (dict(ALL, scaling_lists=True), lambda: vs.with_scaling_lists(gop4(), np.random.default_rng(13))),
    "gop_scaling_lists_3slices_lmcs": (dict(ALL, scaling_lists=True, lmcs=True, **SL3),
                                       lambda: vs.with_lmcs(vs.with_scaling_lists(gop4(), np.random.default_rng(14)), np.random.default_rng(15))),
    "gop_monochrome": (dict(ALL, chroma_format=0, cclm=False, jccr=False), gop4),
    "gop_monochrome_lmcs_alf": (dict(ALL, chroma_format=0, cclm=False, jccr=False, lmcs=True, alf=True),
                                lambda: vs.with_alf(vs.with_lmcs(gop4(), np.random.default_rng(16), chroma=False), np.random.default_rng(17), chroma=False)),
    "gop_alf": (dict(ALL, alf=True), lambda: vs.with_alf(gop4(), np.random.default_rng(3), cc=False)),
    "gop_alf_ccalf": (dict(ALL, alf=True, ccalf=True), lambda: vs.with_alf(gop4(), np.random.default_rng(4))),
    "gop_lmcs": (dict(ALL, lmcs=True), lambda: vs.with_lmcs(gop4(), np.random.default_rng(5))),
    "I_lmcs_dual_tree": (dict(INTRA, lmcs=True, dual_tree=True), lambda: vs.with_lmcs([vs.Pic(0)], np.random.default_rng(6))),
    "low_delay_alf_lmcs": (dict(ALL, alf=True, ccalf=True, lmcs=True), lambda: vs.with_lmcs(vs.with_alf(low_delay(6), np.random.default_rng(7)), np.random.default_rng(8), every=2)),
    "gop_alf_lmcs_8bit": (dict(ALL, alf=True, ccalf=True, lmcs=True, bit_depth=8), lambda: vs.with_lmcs(vs.with_alf(gop4(), np.random.default_rng(9)), np.random.default_rng(10), bit_depth=8)),
}


def _diff(a, b):
    assert len(a) == len(b), (len(a), len(b))
    return [sum(int((x != y).sum()) for x, y in zip(fa, fb)) for fa, fb in zip(a, b)]


@pytest.fixture(scope="module")
def oracle():
    return helpers.load_oracle()


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("seed", [1, 2])
def test_stream_stock_vs_swapped_decoder(name, seed, oracle):
    from tests import stream_util as su
    kw, pics = CASES[name]
    cfg = vs.Config(**kw)
    aus, drawn, nbins = vs.build_stream(cfg, pics(), seed=seed * 7 + len(name))
    assert min(nbins) > 0
    stock = vs.decode(vs.REF_SO, aus)
    assert len(stock) == len(aus) and _diff(drawn, stock) == [0] * len(aus), "the stock reference does not decode the stream to the pictures it was drawn as"
    swapped, log = su.decode_swapped_cpu(aus, oracle)
    assert len(log) == len(aus)                                         # every picture went through the drop-in class
    assert _diff(swapped, stock) == [0] * len(aus)
    assert any(f[0].std() > 1 for f in stock)                           # not a flat picture
    if "scaling_lists" in name: assert all(l["scaling"] > 0 for l in log)            # the tools the case is about reached the work lists
    if "weighted" in name: assert any(l["wp"] > 0 for l in log)
    if "3slices" in name or "_slices_" in name: assert all(l["lfSlices"] == len(cfg.slice_rows) for l in log)


def test_long_stream_recycles_pictures_and_slots(oracle):
    """28 pictures (seven GOPs, open on P anchors): the reference's PicListManager recycles Picture objects and the class's DPB slots turn over
    (17 slots, DecLibReconB200::slotLocked evicts what is no longer referenced)."""
    from tests import stream_util as su
    cfg = vs.Config(**dict(ALL, width=128, height=64))
    pics = []
    for k in range(7): pics += gop4(4 * k, idr=(k == 0))[(0 if k == 0 else 1):]
    aus, drawn, nbins = vs.build_stream(cfg, pics, seed=5)
    stock = vs.decode(vs.REF_SO, aus)
    assert _diff(drawn, stock) == [0] * len(aus)
    swapped, log = su.decode_swapped_cpu(aus, oracle)
    assert _diff(swapped, stock) == [0] * len(aus)
    assert len({l["slot"] for l in log}) < len(log)                     # slots were reused


def test_swapped_decoder_with_a_thread_pool(oracle):
    """the same through DecLib's thread pool (parse and reconstruction tasks on 4 threads, two recon instances in flight)"""
    from tests import stream_util as su
    cfg = vs.Config(**ALL)
    aus, drawn, _ = vs.build_stream(cfg, gop4() + gop4(4, idr=False)[1:], seed=11)
    stock = vs.decode(vs.REF_SO, aus, threads=4)
    assert _diff(drawn, stock) == [0] * len(aus)
    swapped, log = su.decode_swapped_cpu(aus, oracle, threads=4)
    assert _diff(swapped, stock) == [0] * len(aus)


def sequence_change_stream():
    """three coded video sequences in one stream: 256x128 CTU 64 10 bit, 192x192 CTU 32 10 bit, 256x128 CTU 64 8 bit (each starts with its own SPS / PPS and an IDR)"""
    a1, d1, _ = vs.build_stream(vs.Config(**dict(ALL, width=256, height=128)), gop4(), seed=1)
    a2, d2, _ = vs.build_stream(vs.Config(**dict(ALL, width=192, height=192, ctu=32, max_bt_inter=32, max_tt_inter=32)), gop4(), seed=2)
    a3, d3, _ = vs.build_stream(vs.Config(**dict(ALL, width=256, height=128, bit_depth=8)), gop4(), seed=3)
    return a1 + a2 + a3, d1 + d2 + d3


def test_new_sequence_with_another_picture_size_ctu_size_and_bit_depth(oracle):
    """the class finishes what the other instance still holds and rebuilds its device context when an IRAP picture starts a sequence with another geometry"""
    from tests import stream_util as su
    aus, drawn = sequence_change_stream()
    stock = vs.decode(vs.REF_SO, aus, threads=4, frame_samples=256 * 192 * 2)
    assert [f[0].shape for f in stock] == [(128, 256)] * 5 + [(192, 192)] * 5 + [(128, 256)] * 5
    assert _diff(drawn, stock) == [0] * 15
    swapped, log = su.decode_swapped_cpu(aus, oracle, threads=4, frame_samples=256 * 192 * 2)
    assert _diff(swapped, stock) == [0] * 15


@pytest.mark.parametrize("method,bit_depth,threads,structure", [("md5", 10, 1, "gop"), ("crc", 10, 4, "low_delay"), ("checksum", 8, 0, "gop"), ("md5", 8, 2, "low_delay")])
def test_decoded_picture_hash_sei_with_pictures_completed_by_a_pool_task(oracle, method, bit_depth, threads, structure):
    """Every picture carries a decoded-picture-hash SEI (MD5 / CRC / checksum of what it was drawn as); the decoders verify themselves (verifyPictureHash,
    vvdec_get_hash_error_count).  With parseFrameDelay == 0 the parser meets the SEI and waits for pic->reconDone ON THE API THREAD (DecLibParser.cpp:249-261): the
    class has to complete the picture without waitForPrevDecompressedPic() being called — setAsyncFinish — or the decoder never returns."""
    from tests import stream_util as su
    pics = gop4() if structure == "gop" else low_delay(6)
    aus, drawn, _ = vs.build_stream(vs.Config(**dict(ALL, bit_depth=bit_depth)), pics, seed=4, hash_sei=method)
    stock = vs.decode(vs.REF_SO, aus, threads=threads)
    assert vs.decode.hash_errors == 0 and _diff(drawn, stock) == [0] * len(aus)
    damaged = list(aus); b = bytearray(damaged[2]); b[-3] ^= 0x55; damaged[2] = bytes(b)          # (a wrong digest is noticed)
    vs.decode(vs.REF_SO, damaged, threads=threads); assert vs.decode.hash_errors == 1
    swapped, _ = su.decode_swapped_cpu(aus, oracle, threads=threads, async_finish=True)
    assert vs.decode.hash_errors == 0 and _diff(swapped, stock) == [0] * len(aus)


@pytest.mark.parametrize("name", ["gop_all_tools", "low_delay_alf_lmcs", "gop_4tiles_4slices", "gop_3slices_alf_lmcs"])
def test_stream_cases_with_pictures_completed_by_a_pool_task(oracle, name):
    from tests import stream_util as su
    kw, pics = CASES[name]
    aus, drawn, _ = vs.build_stream(vs.Config(**kw), pics(), seed=31)
    stock = vs.decode(vs.REF_SO, aus, threads=4)
    for threads in (1, 4):
        swapped, _ = su.decode_swapped_cpu(aus, oracle, threads=threads, async_finish=True)
        assert _diff(swapped, stock) == [0] * len(aus)


@pytest.mark.parametrize("seed", [2003, 2017, 2130, 3179, 5001, 5014])
def test_random_streams(oracle, seed):
    """a few draws of tools/stream_fuzz.py (random parameter sets / structures / tools; odd seeds: hash SEIs and completion by a pool task) as a regression"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import stream_fuzz as sf
    from tests import stream_util as su
    kw, pics, _ = sf.random_case(seed)
    aus, drawn, _ = vs.build_stream(vs.Config(**kw), pics, seed=seed, hash_sei="md5" if seed & 1 else None)
    stock = vs.decode(vs.REF_SO, aus)
    assert _diff(drawn, stock) == [0] * len(aus)
    swapped, _ = su.decode_swapped_cpu(aus, oracle, threads=1 + 3 * (seed % 3 == 0), async_finish=bool(seed & 1))
    assert _diff(swapped, stock) == [0] * len(aus) and vs.decode.hash_errors == 0


def test_refused_picture_surfaces_as_unsupported(oracle):
    """What the device path leaves to the stock back end is refused on the API thread in decompressPicture (DecLib::reconPicture records it) and comes out of
    vvdec_decode as VVDEC_ERR_NOT_SUPPORTED.  Case: CIIP under LMCS with a 32x32 maximum transform size — the reference maps residual-free CIIP blocks of CUs
    larger than a transform unit through the LMCS forward curve twice (DecCu.cpp:466-476 after CABACReader.cpp:1449-1456 cleared rootCbf); the stock decoder
    still decodes the stream to what it was drawn as."""
    from tests import stream_util as su
    cfg = vs.Config(**dict(ALL, lmcs=True, max_tb64=False))
    aus, drawn, _ = vs.build_stream(cfg, vs.with_lmcs(gop4(), np.random.default_rng(3)), seed=3)
    assert _diff(drawn, vs.decode(vs.REF_SO, aus)) == [0] * len(aus)
    with pytest.raises(vs.DecodeError, match="(?s)unsupported feature.*CIIP under LMCS"):
        su.decode_swapped_cpu(aus, oracle)


@pytest.mark.parametrize("threads", [1, 4])
def test_corrupted_streams_come_back_with_an_error(threads):
    """Bit flips in the slice data: the parser throws while the class's tasks for the picture are scheduled or waiting on the parse barriers.  The error contract of
    DecLibRecon (DecLibRecon.cpp:684-722: first exception parked, tasks drained, picture marked) must bring every stream back — an error code or frames, no crash, no hang."""
    import subprocess, sys
    p = subprocess.run([sys.executable, "-m", "tests.stream_util", "7", "6", str(threads)], capture_output=True, text=True, timeout=600,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    lines = [l for l in p.stdout.splitlines() if l.startswith("ok")]
    assert p.returncode == 0 and len(lines) == 6, (p.returncode, p.stdout[-400:], p.stderr[-400:])
    assert any(l.startswith("ok error") for l in lines)


def test_arithmetic_encoder_round_trip():
    """ref_cabac_encode against the reference's BinDecoder: random context / bypass / terminate sequences come back bin for bin (the generating build
    reads them with drawn bins, the stock build decodes the bytes — compared through a whole slice in the tests above; here: the stop-bit / carry paths
    on many short segments, via streams of one tiny picture)."""
    cfg = vs.Config(width=64, height=64, **INTRA)
    for seed in range(1, 25):
        aus, drawn, nbins = vs.build_stream(cfg, [vs.Pic(0)], seed=seed)
        assert _diff(drawn, vs.decode(vs.REF_SO, aus)) == [0]


def test_emulation_prevention_and_exp_golomb():
    assert vs.escape(bytes([0, 0, 1, 0, 0, 0, 0, 3, 0, 0])) == bytes([0, 0, 3, 1, 0, 0, 3, 0, 0, 3, 3, 0, 0, 3])
    w = vs.Bits().ue(0).ue(1).ue(2).ue(7).se(1).se(-1).se(0)
    assert "".join(map(str, w.b)) == "1" "010" "011" "0001000" "010" "011" "1"
