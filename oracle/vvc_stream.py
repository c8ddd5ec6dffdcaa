"""TEST INFRASTRUCTURE (not shipped, not on the product path): a minimal VVC bitstream WRITER, so that the DecLibRecon seam can be tested from a real
bitstream.

Three steps (see oracle/gen_bindecoder.cpp for why this yields valid streams with every CU-level tool the parameter sets switch on):
  1. the high-level syntax — SPS, PPS, APSs, picture / slice headers — is written here, field by field in the order the reference READS it
     (HLSyntaxReader.cpp: parseSPS :1421, parsePPS :205, parseAPS :855, parsePictureHeader :2694, parseSliceHeader :3438, parseRefPicList :112);
  2. the slice data is DRAWN: oracle/_ref/libvvdec_gen.so (the reference with its BinDecoder replaced by a recording generator) decodes the headers
     followed by a one-byte dummy payload; its CABACReader decides every context / binarisation, the generator decides every bin;
  3. the recorded (context, bin) sequences are arithmetic-encoded (ref_cabac_encode: the reference's probability models, our encoder) and spliced
     behind the slice headers; emulation prevention and start codes make it an Annex-B stream.
The stock reference (oracle/_ref/libvvdec_ref.so: ref_decode_stream) must then decode the stream to exactly the pictures step 2 reconstructed."""
import ctypes as C, os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libvvdec_ref.so")
GEN_SO = os.path.join(HERE, "_ref", "libvvdec_gen.so")
SWAP_SO = os.path.join(HERE, "_ref", "libvvdec_swapped.so")

NAL_TRAIL, NAL_STSA, NAL_RADL, NAL_RASL, NAL_IDR_W_RADL, NAL_IDR_N_LP, NAL_CRA, NAL_GDR = 0, 1, 2, 3, 7, 8, 9, 10
NAL_VPS, NAL_SPS, NAL_PPS, NAL_PREFIX_APS, NAL_SUFFIX_APS, NAL_PH, NAL_AUD = 14, 15, 16, 17, 18, 19, 20
NAL_PREFIX_SEI, NAL_SUFFIX_SEI = 23, 24
SLICE_B, SLICE_P, SLICE_I = 0, 1, 2


class Bits:
    """RBSP bit writer: u(n), ue(v), se(v), flags."""
    def __init__(self): self.b = []
    def u(self, n, v):
        assert 0 <= v < (1 << n) or n == 0, (n, v)
        for i in range(n - 1, -1, -1): self.b.append((v >> i) & 1)
        return self
    def f(self, v): self.b.append(1 if v else 0); return self
    def ue(self, v):
        assert v >= 0
        n = (v + 1).bit_length() - 1
        return self.u(n, 0).u(n + 1, v + 1)
    def se(self, v): return self.ue(2 * v - 1 if v > 0 else -2 * v)
    def aligned(self): return len(self.b) % 8 == 0
    def align_zero(self):
        while not self.aligned(): self.b.append(0)
        return self
    def trailing(self):                                          # rbsp_trailing_bits() / byte_alignment(): a one, then zeros
        self.b.append(1)
        return self.align_zero()
    def bytes(self):
        assert self.aligned()
        return bytes(int("".join(map(str, self.b[i:i + 8])), 2) for i in range(0, len(self.b), 8))


def escape(rbsp):
    """emulation prevention (7.4.2): 00 00 0x -> 00 00 03 0x"""
    out = bytearray(); zeros = 0
    for byte in rbsp:
        if zeros >= 2 and byte <= 3: out.append(3); zeros = 0
        out.append(byte); zeros = zeros + 1 if byte == 0 else 0
    if out and out[-1] == 0: out.append(3)
    return bytes(out)


def nal_unit(nal_type, rbsp, tid=0, layer=0):
    hdr = Bits().f(0).f(0).u(6, layer).u(5, nal_type).u(3, tid + 1).bytes()
    return b"\x00\x00\x00\x01" + hdr + escape(rbsp)


def log2(v):
    assert v > 0 and v & (v - 1) == 0, v
    return v.bit_length() - 1


DEFAULTS = dict(
    width=256, height=128, ctu=64, bit_depth=10, chroma_format=1, min_cb=4, poc_bits=8, level_idc=102, dpb_size=6,
    # partitioning (luma samples): smallest quad-tree leaf, multi-type-tree depth, largest binary / ternary split, per slice class
    min_qt_intra=8, mtt_depth_intra=2, max_bt_intra=32, max_tt_intra=32, dual_tree=False,
    min_qt_intra_c=8, mtt_depth_intra_c=2, max_bt_intra_c=32, max_tt_intra_c=32,            # chroma tree of dual-tree I slices (in luma samples)
    min_qt_inter=8, mtt_depth_inter=2, max_bt_inter=64, max_tt_inter=64,
    max_tb64=True, transform_skip=False, bdpcm=False, mts=False, mts_intra=False, mts_inter=False, lfnst=False, jccr=False,
    sao=False, alf=False, ccalf=False, lmcs=False, weighted_pred=False, weighted_bipred=False,
    temporal_mvp=False, sbtmvp=False, amvr=False, bdof=False, smvd=False, dmvr=False, mmvd=False, max_merge=6, sbt=False, affine=False, affine_6param=False,
    affine_amvr=False, prof=False, max_sub_merge=5, bcw=False, ciip=False, gpm=False, max_gpm=6, isp=False, mrl=False, mip=False, cclm=False,
    chroma_collocated=(True, False), dep_quant=False, sign_hiding=False, scaling_lists=False,
    init_qp=32, cu_qp_delta=False, cabac_init_present=False, deblocking_disabled=False, beta_offset_div2=0, tc_offset_div2=0,
    chroma_qp_tables=None,         # None: one table, identity from 27 up; else 1 (shared), 2 or 3 (with jccr) tables of (start_minus26, [(delta_in_minus1, delta_out), ...])
    ladf=None,                     # luma-adaptive deblocking: (lowest interval QP offset, [(QP offset, delta_threshold_minus1), ...]) with 1..4 further intervals
    ph_tool_control=False,         # sps_bdof / dmvr / prof _control_present_in_ph_flag: pictures switch the tools (Pic.bdof / dmvr / prof)
    lfnst_scaling_disabled=False, min_qp_prime_ts=0, parallel_merge_level=2, ts_max_size=5, cb_cr_deblock_offsets=None,       # (cb beta, cb tc, cr beta, cr tc) with chroma_qp_offsets
    tiles=None,                    # ((column widths), (row heights)) in CTUs; with slice_per_tile every tile is a slice, else one slice holds all tiles
    slice_per_tile=False, lf_across_tiles=True,
    slice_rows=None,               # several rectangular slices in the one tile: CTU rows per slice, e.g. (1, 2, 1); None: one slice, picture header in the slice header
    lf_across_slices=True, deblocking_override=False,                # in-loop filters across slice borders; per-slice deblocking offsets
    chroma_qp_offsets=None,        # (cb, cr, joint) PPS offsets; slices add their own (sh_cb_qp_offset ...) and CUs pick from a list (cu_chroma_qp_offset) if set
    slice_chroma_qp_offsets=False, cu_chroma_qp_offset_list=(),       # list entries: (cb, cr, joint)
)


class Config(dict):
    def __init__(self, **kw):
        bad = set(kw) - set(DEFAULTS); assert not bad, bad
        super().__init__(DEFAULTS); self.update(kw)
    __getattr__ = dict.__getitem__


def write_sps(c):
    """HLSyntaxReader::parseSPS (HLSyntaxReader.cpp:1421-2316), single layer, no sub-pictures, no VUI / HRD."""
    w = Bits()
    ctb_log2, min_cb_log2 = log2(c.ctu), log2(c.min_cb)
    w.u(4, 0).u(4, 0).u(3, 0).u(2, c.chroma_format).u(2, ctb_log2 - 5)
    w.f(1)                                                       # sps_ptl_dpb_hrd_params_present_flag
    w.u(7, 1).f(0).u(8, c.level_idc).f(1).f(0)                   # profile_tier_level(): Main 10, main tier, level, frame only, single layer
    w.f(0).align_zero()                                          #   general_constraints_info(): gci_present_flag = 0, alignment
    w.align_zero().u(8, 0)                                       #   ptl alignment; ptl_num_sub_profiles
    w.f(0).f(0)                                                  # gdr, ref_pic_resampling
    w.ue(c.width).ue(c.height).f(0).f(0)                         # size, no conformance window, no sub-pictures
    w.ue(c.bit_depth - 8).f(0).f(0)                              # bit depth, no wavefronts, no entry points
    w.u(4, c.poc_bits - 4).f(0).u(2, 0).u(2, 0)                  # POC lsb bits, no msb cycle, no extra PH / SH bits
    w.ue(c.dpb_size - 1).ue(min(c.dpb_size - 1, 4)).ue(0)        # dpb_parameters(): max_dec_pic_buffering_minus1, max_num_reorder_pics, max_latency_increase_plus1
    w.ue(min_cb_log2 - 2).f(0)                                   # min CB, no partition-constraint override
    min_qt_i = log2(c.min_qt_intra)
    w.ue(min_qt_i - min_cb_log2).ue(c.mtt_depth_intra)
    if c.mtt_depth_intra: w.ue(log2(c.max_bt_intra) - min_qt_i).ue(log2(c.max_tt_intra) - min_qt_i)
    if c.chroma_format: w.f(c.dual_tree)
    if c.dual_tree:
        min_qt_c = log2(c.min_qt_intra_c)
        w.ue(min_qt_c - min_cb_log2).ue(c.mtt_depth_intra_c)
        if c.mtt_depth_intra_c: w.ue(log2(c.max_bt_intra_c) - min_qt_c).ue(log2(c.max_tt_intra_c) - min_qt_c)
    min_qt_p = log2(c.min_qt_inter)
    w.ue(min_qt_p - min_cb_log2).ue(c.mtt_depth_inter)
    if c.mtt_depth_inter: w.ue(log2(c.max_bt_inter) - min_qt_p).ue(log2(c.max_tt_inter) - min_qt_p)
    if c.ctu > 32: w.f(c.max_tb64)
    w.f(c.transform_skip)
    if c.transform_skip: w.ue(c.ts_max_size - 2).f(c.bdpcm)      # largest transform-skip block (log2)
    w.f(c.mts)
    if c.mts: w.f(c.mts_intra).f(c.mts_inter)
    w.f(c.lfnst)
    if c.chroma_format:
        w.f(c.jccr)
        if c.chroma_qp_tables is None:
            w.f(1).se(0).ue(0).ue(0).ue(0)                       # one chroma QP table for Cb / Cr / joint: start at 26, one pivot (27 -> 26), slope 1 above
        else:
            t = c.chroma_qp_tables; assert len(t) in ((1, 3) if c.jccr else (1, 2))
            w.f(len(t) == 1)                                     # sps_same_qp_table_for_chroma_flag
            for start, pts in t:
                w.se(start).ue(len(pts) - 1)
                for din, dout in pts: w.ue(din).ue(dout ^ din)   # sps_delta_qp_diff_val = delta out XOR delta_in_minus1
    w.f(c.sao).f(c.alf)
    if c.alf and c.chroma_format: w.f(c.ccalf)
    w.f(c.lmcs).f(c.weighted_pred).f(c.weighted_bipred).f(0)     # ..., no long-term references
    w.f(0).f(1).ue(0)                                            # sps_idr_rpl_present_flag, rpl1_same_as_rpl0, no RPL candidates (lists come in the slice headers)
    w.f(0)                                                       # wrap-around
    w.f(c.temporal_mvp)
    if c.temporal_mvp: w.f(c.sbtmvp)
    w.f(c.amvr).f(c.bdof)
    if c.bdof: w.f(c.ph_tool_control)
    w.f(c.smvd).f(c.dmvr)
    if c.dmvr: w.f(c.ph_tool_control)
    w.f(c.mmvd)
    if c.mmvd: w.f(0)
    w.ue(6 - c.max_merge).f(c.sbt).f(c.affine)
    if c.affine:
        w.ue(5 - c.max_sub_merge).f(c.affine_6param)
        if c.amvr: w.f(c.affine_amvr)
        w.f(c.prof)
        if c.prof: w.f(c.ph_tool_control)
    w.f(c.bcw).f(c.ciip)
    if c.max_merge >= 2:
        w.f(c.gpm)
        if c.gpm and c.max_merge >= 3: w.ue(c.max_merge - c.max_gpm)
    w.ue(c.parallel_merge_level - 2)                             # log2 of the parallel merge level
    w.f(c.isp).f(c.mrl).f(c.mip)
    if c.chroma_format: w.f(c.cclm)
    if c.chroma_format == 1: w.f(c.chroma_collocated[0]).f(c.chroma_collocated[1])
    w.f(0)                                                       # palette
    if c.chroma_format == 3 and not c.max_tb64: w.f(0)           # ACT
    if c.transform_skip: w.ue(c.min_qp_prime_ts)                 # (sps_internal_bit_depth_minus_input_bit_depth: MinQpPrimeTs = 4 + 6 * value)
    w.f(0).f(c.ladf is not None)                                 # IBC, LADF
    if c.ladf is not None:
        low, steps = c.ladf; assert 1 <= len(steps) <= 4
        w.u(2, len(steps) - 1).se(low)
        for off, thr in steps: w.se(off).ue(thr)
    w.f(c.scaling_lists)
    if c.lfnst and c.scaling_lists: w.f(c.lfnst_scaling_disabled)
    w.f(c.dep_quant).f(c.sign_hiding).f(0)                       # ..., no virtual boundaries
    w.f(0)                                                       # sps_timing_hrd_params_present_flag
    w.f(0).f(0).f(0)                                             # field_seq, VUI, extension
    return nal_unit(NAL_SPS, w.trailing().bytes())


def write_pps(c):
    """HLSyntaxReader::parsePPS (HLSyntaxReader.cpp:205-850): one tile; one slice (pps_no_pic_partition_flag) or rectangular slices of whole CTU rows."""
    w = Bits()
    w.u(6, 0).u(4, 0).f(0).ue(c.width).ue(c.height).f(0).f(0).f(0)      # ids, mixed NAL types, size, conformance / scaling window, output flag
    multi = c.slice_rows is not None
    assert not (multi and c.tiles), "slice_rows (slices inside the one tile) and tiles are alternatives"
    w.f(not (multi or c.tiles)).f(0)                                   # no_pic_partition, no sub-picture id mapping
    wc, hc = -(-c.width // c.ctu), -(-c.height // c.ctu)
    if multi:
        rows = list(c.slice_rows); n = len(rows)
        assert n >= 2 and sum(rows) == hc and min(rows) >= 1 and rows[-1] <= rows[-2], "slice_rows: whole CTU rows, the last slice not taller than the one before"
        w.u(2, log2(c.ctu) - 5).ue(0).ue(0).ue(wc - 1).ue(hc - 1)        # one explicit tile column / row spanning the picture
        w.f(0)                                                         # (one tile: rectangular slices inferred) pps_single_slice_per_subpic_flag
        w.ue(n - 1)
        if n - 1 > 1: w.f(0)                                           # pps_tile_idx_delta_present_flag
        w.ue(n - 1)                                                    # pps_num_exp_slices_in_tile: all but the last height explicit, the rest of the tile is the last slice
        for r in rows[:-1]: w.ue(r - 1)
        w.f(c.lf_across_slices)
    elif c.tiles:
        cols, rows = c.tiles; nt = len(cols) * len(rows)
        assert sum(cols) == wc and sum(rows) == hc and nt > 1
        w.u(2, log2(c.ctu) - 5).ue(len(cols) - 1).ue(len(rows) - 1)
        for v in cols: w.ue(v - 1)
        for v in rows: w.ue(v - 1)
        w.f(c.lf_across_tiles).f(1)                                    # ..., pps_rect_slice_flag
        w.f(not c.slice_per_tile)                                      # pps_single_slice_per_subpic_flag: the picture is the one sub-picture
        if c.slice_per_tile:
            w.ue(nt - 1)
            if nt - 1 > 1: w.f(0)                                      # pps_tile_idx_delta_present_flag
            for t in range(nt - 1):                                    # slice i = tile i: one tile wide and high
                if t % len(cols) != len(cols) - 1: w.ue(0)
                if t // len(cols) != len(rows) - 1 and t % len(cols) == 0: w.ue(0)
                if rows[t // len(cols)] > 1: w.ue(0)                   # pps_num_exp_slices_in_tile: the tile is one slice
        w.f(c.lf_across_slices)
    w.f(c.cabac_init_present).ue(0).ue(0).f(0)                         # one active reference per list by default, no rpl1 index
    w.f(c.weighted_pred).f(c.weighted_bipred).f(0)                     # ..., wrap-around
    w.se(c.init_qp - 26).f(c.cu_qp_delta)
    w.f(c.chroma_qp_offsets is not None)                               # pps_chroma_tool_offsets_present_flag
    if c.chroma_qp_offsets is not None:
        cb, cr, joint = c.chroma_qp_offsets
        w.se(cb).se(cr).f(c.jccr)
        if c.jccr: w.se(joint)
        w.f(c.slice_chroma_qp_offsets).f(len(c.cu_chroma_qp_offset_list) > 0)
        if c.cu_chroma_qp_offset_list:
            w.ue(len(c.cu_chroma_qp_offset_list) - 1)
            for e in c.cu_chroma_qp_offset_list:
                w.se(e[0]).se(e[1])
                if c.jccr: w.se(e[2])
    w.f(1).f(c.deblocking_override).f(c.deblocking_disabled)           # deblocking control present, override enabled, disabled flag
    if (multi or c.tiles) and c.deblocking_override: w.f(0)            # pps_dbf_info_in_ph_flag
    if not c.deblocking_disabled:
        w.se(c.beta_offset_div2).se(c.tc_offset_div2)
        if c.chroma_qp_offsets is not None:
            for v in (c.cb_cr_deblock_offsets or (c.beta_offset_div2, c.tc_offset_div2) * 2): w.se(v)
    if multi or c.tiles: w.f(0).f(0).f(0).f(0)                         # RPL / SAO / ALF / QP delta stay in the slice headers (no weighted-prediction tables in the PH either)
    w.f(0).f(0).f(0)                                                   # PH / SH extension, PPS extension
    return nal_unit(NAL_PPS, w.trailing().bytes())


ALF_APS, LMCS_APS, SCALING_LIST_APS = 0, 1, 2


def write_aps(aps_type, aps_id, body, chroma_present=True):
    """HLSyntaxReader::parseAPS (HLSyntaxReader.cpp:855-903): type, id, chroma flag, payload, no extension."""
    w = Bits().u(3, aps_type).u(5, aps_id).f(chroma_present)
    body(w)
    return nal_unit(NAL_PREFIX_APS, w.f(0).trailing().bytes())


def random_alf_aps(rng, luma=True, chroma=True, cc=(True, True)):
    """ALF APS content: luma classes -> up to 25 filters of 12 coded taps (+ clipping indices), up to 8 chroma alternatives of 6 taps, up to 4 cross-component
    filters of 7 power-of-two taps per chroma component.  Tap magnitudes stay small enough for the sum constraints of the filters' centre taps."""
    a = dict(luma=None, chroma=None, cc=[None, None])
    if luma:
        n = int(rng.integers(1, 26))
        a["luma"] = dict(clip=bool(rng.integers(0, 2)), n=n, delta_idx=[int(rng.integers(0, n)) for _ in range(25)] if n > 1 else [0] * 25,
                         coeff=rng.integers(-12, 13, size=(n, 12)).tolist(), clip_idx=rng.integers(0, 4, size=(n, 12)).tolist())
    if chroma:
        n = int(rng.integers(1, 9))
        a["chroma"] = dict(clip=bool(rng.integers(0, 2)), n=n, coeff=rng.integers(-20, 21, size=(n, 6)).tolist(), clip_idx=rng.integers(0, 4, size=(n, 6)).tolist())
    for k in range(2):
        if cc[k]:
            n = int(rng.integers(1, 5))
            a["cc"][k] = dict(n=n, mapped=rng.integers(0, 5, size=(n, 7)).tolist(), sign=rng.integers(0, 2, size=(n, 7)).tolist())
    return a


def write_alf_aps(aps_id, a, chroma_present=True):
    """parseAlfAps / alfFilterCoeffs (HLSyntaxReader.cpp:905-1012, 4659-4710)"""
    def coeffs(w, n, taps, f):
        for i in range(n):
            for j in range(taps):
                v = f["coeff"][i][j]; w.ue(abs(v))
                if v: w.f(v < 0)
        if f["clip"]:
            for i in range(n):
                for j in range(taps): w.u(2, f["clip_idx"][i][j])
    def body(w):
        w.f(a["luma"] is not None)
        if chroma_present: w.f(a["chroma"] is not None).f(a["cc"][0] is not None).f(a["cc"][1] is not None)
        else: assert a["chroma"] is None and a["cc"] == [None, None]
        if a["luma"]:
            f = a["luma"]; w.f(f["clip"]).ue(f["n"] - 1)
            if f["n"] > 1:
                bits = (f["n"] - 1).bit_length()
                for k in range(25): w.u(bits, f["delta_idx"][k])
            coeffs(w, f["n"], 12, f)
        if a["chroma"]:
            f = a["chroma"]; w.f(f["clip"]).ue(f["n"] - 1)
            for alt in range(f["n"]): coeffs(w, 1, 6, dict(coeff=[f["coeff"][alt]], clip=f["clip"], clip_idx=[f["clip_idx"][alt]]))
        for k in range(2):
            f = a["cc"][k]
            if f:
                w.ue(f["n"] - 1)
                for i in range(f["n"]):
                    for j in range(7):
                        w.u(3, f["mapped"][i][j])
                        if f["mapped"][i][j]: w.f(f["sign"][i][j])
    return write_aps(ALF_APS, aps_id, body, chroma_present)


def random_lmcs_aps(rng, bit_depth=10):
    """LMCS model: bins [min, max] with codeword deltas around OrgCW = 2^bitDepth / 16, the sum of the codewords below 2^bitDepth"""
    org = (1 << bit_depth) // 16
    lo, hi = int(rng.integers(0, 3)), int(rng.integers(12, 16))
    while True:
        d = rng.integers(-org // 2, org // 2 + 1, size=hi - lo + 1)
        if (org * len(d) + d.sum()) <= (1 << bit_depth) - 1: break
    crs_min = max(-7, (org >> 3) - int(org + d.min()))           # lmcsCW[i] + lmcsDeltaCrs stays within [OrgCW >> 3, (OrgCW << 3) - 1] (Reshape.cpp:335)
    return dict(min_bin=lo, max_bin=hi, delta=d.tolist(), crs=int(rng.integers(crs_min, 8)))


def write_lmcs_aps(aps_id, m, chroma_present=True):
    """parseLmcsAps (HLSyntaxReader.cpp:1014-1054)"""
    def body(w):
        prec = max(1, max(abs(v) for v in m["delta"]).bit_length())
        w.ue(m["min_bin"]).ue(15 - m["max_bin"]).ue(prec - 1)
        for v in m["delta"]:
            w.u(prec, abs(v))
            if v: w.f(v < 0)
        if chroma_present:
            w.u(3, abs(m["crs"]))
            if m["crs"]: w.f(m["crs"] < 0)
    return write_aps(LMCS_APS, aps_id, body, chroma_present)


def _diag_scan8():
    """the 8x8 up-right diagonal scan as (x, y) pairs (g_scanOrder[SCAN_UNGROUPED]: anti-diagonals from the bottom-left end)"""
    out = []
    for d in range(15):
        for y in range(min(d, 7), -1, -1):
            x = d - y
            if x < 8: out.append((x, y))
    return out


def random_scaling_list_aps(rng):
    """28 lists (2 of 2x2, 6 of 4x4, 20 of 8x8 with a DC value from the 16x16 lists on): each either copied (from the flat 16 default or an earlier list) or
    coded explicitly as DPCM of values around 16"""
    lists = []
    for i in range(28):
        n = 2 if i < 2 else 4 if i < 8 else 8
        max_delta = i if i < 2 else (i - 2 if i < 8 else i - 8)
        if rng.random() < 0.35:
            lists.append(dict(copy=True, pred_delta=int(rng.integers(0, max_delta + 1)) if i not in (0, 2, 8) else 0))
        else:
            walk = np.clip(16 + np.cumsum(rng.integers(-3, 4, size=n * n + 1)), 4, 120)
            lists.append(dict(copy=False, dc=int(walk[0]), values=[int(v) for v in walk[1:]]))
    return lists


def write_scaling_list_aps(aps_id, lists, chroma_present=True):
    """parseScalingList / decodeScalingList (HLSyntaxReader.cpp:4509-4628): copy mode or explicit DPCM (no prediction from another list + deltas)"""
    scan = _diag_scan8()
    def body(w):
        for i, l in enumerate(lists):
            if not chroma_present and i % 3 != 2 and i != 27: continue      # (luma lists only: ScalingList::isLumaScalingList)
            w.f(l["copy"])
            if l["copy"]:
                if i not in (0, 2, 8): w.ue(l["pred_delta"])
                continue
            w.f(0)                                                          # scaling_list_pred_mode_flag
            n = 2 if i < 2 else 4 if i < 8 else 8
            nxt = 0
            if i > 13:
                w.se(l["dc"] - 8); nxt = l["dc"] - 8
            for k in range(n * n):
                if i > 25 and scan[k][0] >= 4 and scan[k][1] >= 4: continue
                d = (l["values"][k] - 8) - nxt
                w.se(d); nxt += d
    return write_aps(SCALING_LIST_APS, aps_id, body, chroma_present)


class Pic(dict):
    """One picture of the stream: poc, slice_type, refs = ([POCs list 0], [POCs list 1]), qp, tid and per-picture tool switches."""
    def __init__(self, poc, slice_type=SLICE_I, refs=((), ()), qp=None, idr=None, referenced=True, **kw):
        super().__init__(poc=poc, slice_type=slice_type, refs=refs, qp=qp, idr=(poc == 0 if idr is None else idr), referenced=referenced,
                         sao=(True, True), dep_quant=True, sign_hiding=True, temporal_mvp=True, col_from_l0=True, mvd_l1_zero=False, cabac_init=False,
                         aps=[],            # parameter-set NAL units (write_alf_aps / write_lmcs_aps) sent ahead of this picture
                         alf=None,          # dict(luma=[APS ids], cb=bool, cr=bool, chroma_aps=id, cc_cb=id or None, cc_cr=id or None)
                         lmcs=None,         # dict(aps=id, chroma_scale=bool)
                         slice_types=None,  # several slices: a type per slice (default: the picture's); I slices may sit in P / B pictures
                         wp=None,           # seed of the explicit prediction weights (streams with weighted_pred / weighted_bipred)
                         scaling_list=None, # id of the scaling-list APS the picture quantises with
                         bdof=True, dmvr=True, prof=True, jccr_sign=False)    # picture-level tool switches (ph_tool_control), sign of the joint Cb-Cr residual
        bad = set(kw) - set(self); assert not bad, bad
        self.update(kw)
    __getattr__ = dict.__getitem__


def write_ref_pic_list(w, c, poc, ref_pocs):
    """parseRefPicList (HLSyntaxReader.cpp:112-200) for a list signalled in the header (rplsIdx = -1): short-term entries only."""
    w.ue(len(ref_pocs))
    prev = 0
    for i, r in enumerate(ref_pocs):
        delta = (r - poc) - prev                                   # the deltas accumulate: entry i refers to POC poc + sum( delta[0..i] ) (Slice.cpp:484)
        prev += delta
        a = abs(delta)
        if (not c.weighted_pred and not c.weighted_bipred) or i == 0:
            assert a >= 1, "equal POCs need abs_delta_poc_st = 0, which only later entries of weighted-prediction streams can signal"
            w.ue(a - 1)
        else:
            w.ue(a)
        if a > 0: w.f(delta < 0)


def write_pred_weight_table(w, c, rng, n_active, slice_type):
    """parsePredWeightTable (HLSyntaxReader.cpp:4359-4508), tables in the slice header: per list luma flags, chroma flags, then weights / offsets"""
    denom = int(rng.integers(4, 8)); w.ue(denom)
    if c.chroma_format: w.se(int(rng.integers(-1, 2)) if denom < 7 else -1)
    for l in (0, 1):
        n = n_active[l] if (l == 0 or (slice_type == SLICE_B and c.weighted_bipred)) else 0
        luma = [bool(rng.integers(0, 2)) for _ in range(n)]; chroma = [bool(rng.integers(0, 2)) for _ in range(n)]
        for f in luma: w.f(f)
        if c.chroma_format:
            for f in chroma: w.f(f)
        for i in range(n):
            if luma[i]: w.se(int(rng.integers(-10, 11))).se(int(rng.integers(-20, 21)))
            if c.chroma_format and chroma[i]:
                for _ in range(2): w.se(int(rng.integers(-10, 11))).se(int(rng.integers(-30, 31)))


def write_picture_header(w, c, p):
    """picture_header_structure(): parsePictureHeader (HLSyntaxReader.cpp:2694-3360)"""
    types = slice_types_of(c, p)
    inter, intra = any(t != SLICE_I for t in types), any(t == SLICE_I for t in types)
    w.f(p.idr).f(not p.referenced)
    if p.idr: w.f(0)                                             # ph_gdr_pic_flag
    w.f(inter)
    if inter: w.f(intra)                                         # ph_intra_slice_allowed_flag
    w.ue(0).u(c.poc_bits, p.poc & ((1 << c.poc_bits) - 1))
    if c.lmcs:
        w.f(p.lmcs is not None)                                  # ph_lmcs_enabled_flag
        if p.lmcs is not None:
            w.u(2, p.lmcs["aps"])
            if c.chroma_format: w.f(p.lmcs.get("chroma_scale", True))
    if c.scaling_lists:
        w.f(p.scaling_list is not None)                          # ph_explicit_scaling_list_enabled_flag
        if p.scaling_list is not None: w.u(3, p.scaling_list)
    if intra or not inter:
        if c.cu_qp_delta: w.ue(0)                                # ph_cu_qp_delta_subdiv_intra_slice
        if c.cu_chroma_qp_offset_list: w.ue(0)
    if inter:
        if c.cu_qp_delta: w.ue(0)
        if c.cu_chroma_qp_offset_list: w.ue(0)
        if c.temporal_mvp: w.f(p.temporal_mvp)
        w.f(p.mvd_l1_zero)                                       # RPLs are in the slice headers: ph_mvd_l1_zero_flag is always present
        if c.bdof and c.ph_tool_control: w.f(not p.bdof)         # ph_bdof_disabled_flag
        if c.dmvr and c.ph_tool_control: w.f(not p.dmvr)
        if c.affine and c.prof and c.ph_tool_control: w.f(not p.prof)
    if c.jccr: w.f(p.jccr_sign)                                  # ph_joint_cbcr_sign_flag


def slice_types_of(c, p):
    n = len(c.slice_rows) if c.slice_rows is not None else len(c.tiles[0]) * len(c.tiles[1]) if (c.tiles and c.slice_per_tile) else 1
    return list(p.slice_types) if p.slice_types is not None else [p.slice_type] * n


def write_slices(c, p):
    """The VCL side of one picture: [(nal type, rbsp bytes up to and including byte_alignment())] — a PH NAL first when the picture has several slices.
    parseSliceHeader (HLSyntaxReader.cpp:3438-4065)."""
    import numpy as _np
    types = slice_types_of(c, p)
    multi = len(types) > 1
    irap = p.idr
    nals = []
    segments = 1 if (not c.tiles or c.slice_per_tile) else len(c.tiles[0]) * len(c.tiles[1])      # CABAC segments per slice: one per tile it holds
    if multi:
        w = Bits(); write_picture_header(w, c, p); nals.append((NAL_PH, w.trailing().bytes(), 0))
    inter_allowed = any(t != SLICE_I for t in types)
    for k, st in enumerate(types):
        inter = st != SLICE_I
        w = Bits()
        w.f(not multi)                                           # sh_picture_header_in_slice_header_flag
        if not multi: write_picture_header(w, c, p)
        if multi and len(types) > 1: w.u((len(types) - 1).bit_length(), k)     # sh_slice_address: index of the slice in the (only) sub-picture
        if inter_allowed: w.ue(st)
        if irap: w.f(0)                                          # sh_no_output_of_prior_pics_flag
        if c.alf:
            a = p.alf[k % len(p.alf)] if isinstance(p.alf, (list, tuple)) else p.alf
            w.f(a is not None)                                   # sh_alf_enabled_flag
            if a is not None:
                w.u(3, len(a["luma"]))
                for i in a["luma"]: w.u(3, i)
                if c.chroma_format:
                    w.f(a.get("cb", False)).f(a.get("cr", False))
                    if a.get("cb") or a.get("cr"): w.u(3, a["chroma_aps"])
                if c.ccalf:
                    for key in ("cc_cb", "cc_cr"):
                        w.f(a.get(key) is not None)
                        if a.get(key) is not None: w.u(3, a[key])
        if multi and c.lmcs and p.lmcs is not None: w.f(1)       # sh_lmcs_used_flag
        if multi and c.scaling_lists and p.scaling_list is not None: w.f(1)     # sh_explicit_scaling_list_used_flag
        refs = [list(p.refs[0]), list(p.refs[1])]
        if k & 1: refs = [r[::-1] for r in refs]                 # odd slices list the same pictures in the opposite order
        if not irap:                                             # IDR without sps_idr_rpl_present_flag carries no lists
            for l in (0, 1): write_ref_pic_list(w, c, p.poc, refs[l])
        if inter:
            n0, n1 = len(refs[0]), len(refs[1])
            assert n0 >= 1 and (st != SLICE_B or n1 >= 1)
            if n0 > 1 or (st == SLICE_B and n1 > 1):
                w.f(1)                                           # sh_num_ref_idx_active_override_flag: all entries active
                if n0 > 1: w.ue(n0 - 1)
                if st == SLICE_B and n1 > 1: w.ue(n1 - 1)
            n_active = [n0, n1 if st == SLICE_B else 0]
            if c.cabac_init_present: w.f(p.cabac_init)
            if c.temporal_mvp and p.temporal_mvp:
                if st == SLICE_B: w.f(p.col_from_l0)
                col_l0 = p.col_from_l0 or st != SLICE_B
                if (col_l0 and n_active[0] > 1) or (not col_l0 and n_active[1] > 1): w.ue(0)
            if (c.weighted_pred and st == SLICE_P) or (c.weighted_bipred and st == SLICE_B):
                write_pred_weight_table(w, c, _np.random.default_rng((p.wp or 0) * 131 + k), n_active, st)
        qp = (c.init_qp if p.qp is None else p.qp) + (k % 3 - (1 if k else 0)) * 2 * (1 if multi else 0)
        w.se(qp - c.init_qp)
        if c.chroma_qp_offsets is not None and c.slice_chroma_qp_offsets:
            w.se((k % 3) - 1).se(1 - (k % 3))
            if c.jccr: w.se(k % 2)
        if c.cu_chroma_qp_offset_list: w.f(1)                    # sh_cu_chroma_qp_offset_enabled_flag
        if c.sao:
            sao = p.sao if not multi else (p.sao[0] and k % 3 != 2, p.sao[1] and k % 4 != 3)
            w.f(sao[0])
            if c.chroma_format: w.f(sao[1])
        if c.deblocking_override:
            w.f(1)                                               # sh_deblocking_params_present_flag
            if not c.deblocking_disabled: w.f(0)                 # sh_deblocking_filter_disabled_flag
            w.se((k * 5 + p.poc) % 7 - 3).se((k * 3 + p.poc) % 5 - 2)
            if c.chroma_qp_offsets is not None: w.se(k % 3 - 1).se(1 - k % 3).se((k + 1) % 3 - 1).se(k % 2)
        dq = c.dep_quant and p.dep_quant and not (multi and k % 2)
        if c.dep_quant: w.f(dq)
        sdh = c.sign_hiding and not dq and p.sign_hiding
        if c.sign_hiding and not dq: w.f(sdh)
        if c.transform_skip and not dq and not sdh: w.f(0)       # sh_ts_residual_coding_disabled_flag
        w.trailing()                                             # byte_alignment()
        nals.append((NAL_IDR_N_LP if irap else NAL_TRAIL, w.bytes(), segments))
    return nals


def with_alf(pics, rng, cc=True, chroma=True):
    """every picture sends a new ALF APS (ids cycle through 0..7) and filters with it and the one before it; CC-ALF filters from both"""
    for i, p in enumerate(pics):
        p["aps"] = p["aps"] + [write_alf_aps(i % 8, random_alf_aps(rng, chroma=chroma, cc=(cc and chroma, cc and chroma)), chroma)]
        if not chroma:
            p["alf"] = dict(luma=[i % 8] + ([(i - 1) % 8] if i else [])); continue
        p["alf"] = dict(luma=[i % 8] + ([(i - 1) % 8] if i else []), cb=True, cr=bool(i & 1) or i == 0, chroma_aps=i % 8,
                        cc_cb=(i % 8 if cc else None), cc_cr=((i - 1) % 8 if cc and i else None))
    return pics


def with_scaling_lists(pics, rng, chroma_present=True):
    """every other picture sends a new scaling-list APS (ids cycle through 0..7); all pictures quantise with the latest"""
    last = None
    for i, p in enumerate(pics):
        if i % 2 == 0:
            last = (i // 2) % 8
            p["aps"] = p["aps"] + [write_scaling_list_aps(last, random_scaling_list_aps(rng), chroma_present)]
        p["scaling_list"] = last
    return pics


def with_lmcs(pics, rng, bit_depth=10, every=1, chroma=True):
    """pictures send a new LMCS APS (ids cycle through 0..3) and use it, with or without chroma residual scaling"""
    for i, p in enumerate(pics):
        if i % every: continue
        p["aps"] = p["aps"] + [write_lmcs_aps(i % 4, random_lmcs_aps(rng, bit_depth), chroma)]
        p["lmcs"] = dict(aps=i % 4, chroma_scale=bool(i & 1) or i == 0)
    return pics


def write_hash_sei(method, planes, bit_depth):
    """Suffix SEI with the decoded picture hash of `planes` (SEIread.cpp:443 xParseSEIDecodedPictureHash; PicYuvMD5.cpp: MD5 of the little-endian samples per
    component, or the 16-bit CRC / 32-bit checksum restated in oracle/k7_output.c)."""
    import hashlib
    kind = {"md5": 0, "crc": 1, "checksum": 2}[method]
    digest = b""
    for pl in planes:
        if pl.size == 0: continue
        if kind == 0: digest += hashlib.md5((pl.astype("<u2") if bit_depth > 8 else pl.astype("u1")).tobytes()).digest()
        else:
            orc = C.CDLL(os.path.join(HERE, "liboracle.so")); out = np.zeros(16, np.uint8); a = np.ascontiguousarray(pl)
            orc.orc_plane_hash.argtypes = [C.c_int, C.c_int, np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS"), C.c_ssize_t, C.c_int, C.c_int, np.ctypeslib.ndpointer(np.uint8)]
            n = orc.orc_plane_hash(kind, bit_depth, a, a.shape[1], a.shape[1], a.shape[0], out); digest += bytes(out[:n])
    single = len([pl for pl in planes if pl.size]) == 1
    payload = bytes([kind, 0x80 if single else 0]) + digest
    return nal_unit(NAL_SUFFIX_SEI, bytes([132, len(payload)]) + payload + b"\x80")


def output_order(pics):
    """index of every picture (decoding order) among the output frames: IDR periods in turn, POC order inside"""
    groups, cur = [], []
    for i, p in enumerate(pics):
        if p.idr and cur: groups.append(cur); cur = []
        cur.append(i)
    groups.append(cur)
    order = [i for g in groups for i in sorted(g, key=lambda k: pics[k].poc)]
    return {i: n for n, i in enumerate(order)}


# ---- libraries -----------------------------------------------------------------------------------------------------------------------------------------
_libs = {}


def _lib(path):
    if path not in _libs:
        lib = C.CDLL(path)
        u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        lib.ref_decode_stream.argtypes = [u8p, np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS"), C.c_int, C.c_int,
                                          np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS"), C.c_long, C.POINTER(C.c_int), C.c_char_p, C.c_int]
        lib.ref_decode_stream2.argtypes = lib.ref_decode_stream.argtypes + [np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS"), C.c_int]
        lib.ref_cabac_encode.argtypes = [C.c_int, C.c_int, np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS"), u8p, C.c_long, u8p, C.c_long]
        lib.ref_cabac_encode.restype = C.c_long
        lib.ref_ctx_offset.argtypes = [C.c_char_p]
        lib.ref_last_hash_errors.restype = C.c_int
        lib.ref_ctx_names.restype = C.c_char_p
        if hasattr(lib, "gen_segment"):
            lib.gen_segment.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_long]; lib.gen_segment.restype = C.c_long
            lib.gen_reset.argtypes = [C.c_uint32]; lib.gen_set_bias.argtypes = [C.c_int] * 3
        _libs[path] = lib
    return _libs[path]


def available():
    return os.path.exists(REF_SO) and os.path.exists(GEN_SO)


class DecodeError(RuntimeError):
    pass


def decode(lib_path, aus, threads=1, max_frames=None, frame_samples=None):
    """The stream (list of access-unit byte strings) through vvdec_decode / vvdec_flush of the given build.  Returns [(Y, Cb, Cr)] in output order."""
    lib = _lib(lib_path)
    stream = np.frombuffer(b"".join(aus), np.uint8).copy()
    offs = np.zeros(len(aus) + 1, np.int64); offs[1:] = np.cumsum([len(a) for a in aus])
    cap = (frame_samples or 1 << 22) * (max_frames or len(aus))
    out = np.zeros(cap, np.int16); dims = (C.c_int * 6)(); err = C.create_string_buffer(1024)
    nmax = (max_frames or len(aus)) + 2; fd = np.zeros(4 * nmax, np.int32)
    n = lib.ref_decode_stream2(stream, offs, len(aus), threads, out, cap, dims, err, len(err), fd, nmax)
    if n < 0: raise DecodeError(f"vvdec error {n}: {err.value.decode(errors='replace')}")
    frames, pos = [], 0
    for i in range(n):
        w, h, cw, ch = (int(v) for v in fd[4 * i:4 * i + 4])
        y = out[pos:pos + w * h].reshape(h, w).copy(); pos += w * h
        cb = out[pos:pos + cw * ch].reshape(ch, cw).copy(); pos += cw * ch
        cr = out[pos:pos + cw * ch].reshape(ch, cw).copy(); pos += cw * ch
        frames.append((y, cb, cr))
    decode.hash_errors = lib.ref_last_hash_errors()              # pictures whose decoded-picture-hash SEI did not match (0 if the stream carries none)
    return frames


def ctx_sets():
    """{name: (offset, size)} of the reference's context sets (Contexts.h ContextSetCfg); arrays of sets as name0, name1, ..."""
    lib = _lib(REF_SO)
    names = []
    for ln in lib.ref_ctx_names().decode().split():
        if ":" in ln:
            n, k = ln.split(":"); names += [f"{n}{i}" for i in range(int(k))]
        else: names.append(ln)
    offs = {n: lib.ref_ctx_offset(n.encode()) for n in names}
    total = lib.ref_ctx_offset(b"total")
    order = sorted(set(offs.values())) + [total]
    return {n: (o, order[order.index(o) + 1] - o) for n, o in offs.items()}


# bin statistics that keep drawn pictures varied but bounded: P( bin = 1 ) * 256 per context set (every other context: 128)
DEFAULT_BIAS = {"SplitFlag": 150, "SplitQtFlag": 140, "QtCbf0": 150, "QtCbf1": 90, "QtCbf2": 90, "SkipFlag": 60, "MergeFlag": 120,
                "SigCoeffGroup0": 90, "SigCoeffGroup1": 90, "DeltaQP": 60, "LastX0": 110, "LastY0": 110, "LastX1": 100, "LastY1": 100,
                "GtxFlag0": 90, "GtxFlag1": 90, "GtxFlag2": 90, "GtxFlag3": 90, "Mvd": 150}
for _k in range(6): DEFAULT_BIAS[f"SigFlag{_k}"] = 90


def build_stream(cfg, pics, seed=1, bias=None, bypass_p=128, max_bypass_run=12, hash_sei=None):
    """Writes the stream for `pics` (decoding order).  Returns (access units, frames the generating library reconstructed, bins per slice).
    hash_sei: "md5" / "crc" / "checksum" — every picture carries a decoded-picture-hash SEI of what it was drawn as (a decoder with verifyPictureHash checks itself)."""
    gen, ref = _lib(GEN_SO), _lib(REF_SO)
    params = write_sps(cfg) + write_pps(cfg)
    heads = [write_slices(cfg, p) for p in pics]                  # per picture: [(nal type, header bytes, is a slice)]
    # step 2: draw the slice data
    gen.gen_reset(seed)
    sets = ctx_sets()
    gen.gen_set_bias(0, 1023, 128)
    for name, p in {**DEFAULT_BIAS, **(bias or {})}.items():
        o, n = sets[name]; gen.gen_set_bias(o, n, p)
    gen.gen_set_bias(-1, 0, bypass_p); gen.gen_set_max_bypass_run(max_bypass_run)
    pre = [(params if i == 0 else b"") + b"".join(p.aps) for i, p in enumerate(pics)]
    aus = [pre[i] + b"".join(nal_unit(t, h + (b"\x80" if vcl else b"")) for t, h, vcl in nals) for i, nals in enumerate(heads)]
    drawn = decode(GEN_SO, aus, frame_samples=cfg.width * cfg.height * 3 // 2 + 64)
    nseg = gen.gen_num_segments()
    assert nseg == sum(vcl for nals in heads for _, _, vcl in nals), (nseg, len(pics))      # (vcl: CABAC segments of the NAL, 0 for a PH)
    # step 3: encode and splice
    out, nbins, seg = [], [], 0
    for i, nals in enumerate(heads):
        au = pre[i]
        for t, h, vcl in nals:
            if not vcl: au += nal_unit(t, h); continue
            data = b""
            for _ in range(vcl):                                  # a tile ends with its own terminating bin, stop bit and alignment; the next starts on fresh contexts
                info = (C.c_int * 3)(); n = gen.gen_segment(seg, info, None, None, 0)
                ctx = np.zeros(n, np.int16); bins = np.zeros(n, np.uint8)
                gen.gen_segment(seg, info, ctx.ctypes.data, bins.ctypes.data, n)
                buf = np.zeros(n // 4 + 64, np.uint8)
                m = ref.ref_cabac_encode(info[0], info[1], ctx, bins, n, buf, len(buf))
                assert m > 0
                data += bytes(buf[:m]); nbins.append(n); seg += 1
            au += nal_unit(t, h + data)
        if hash_sei: au += write_hash_sei(hash_sei, drawn[output_order(pics)[i]], cfg.bit_depth)
        out.append(au)
    return out, drawn, nbins
