"""Randomised bitstream fuzz of the DecLibRecon seam on the CPU (test infrastructure): random parameter sets / GOP structures / tool switches through
oracle/vvc_stream.py, decoded by the stock reference and by the swapped build with the oracle chain as the device (tests/stream_util.py).
    python tools/stream_fuzz.py FIRST_SEED COUNT [big]
Prints one line per stream; exit code 1 if any stream differed."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vvc_stream as vs
from tests import helpers, stream_util as su
from tests.test_stream_cpu import INTRA, INTER, gop4, gop8, low_delay, _diff


def random_case(seed, big=False):
    """big: pictures of 4..9 CTUs a side (up to 1152x1152) instead of 2..4"""
    r = np.random.default_rng(seed)
    pick = lambda *a: a[int(r.integers(len(a)))]
    kw = {k: bool(r.integers(0, 4)) for k in list(INTRA) + list(INTER)}          # each tool on with probability 3/4
    kw["mts_intra"] &= kw["mts"]; kw["mts_inter"] &= kw["mts"]; kw["sbtmvp"] &= kw["temporal_mvp"]
    kw["affine_6param"] &= kw["affine"]; kw["prof"] &= kw["affine"]; kw["affine_amvr"] &= kw["affine"] and kw["amvr"]
    ctu = pick(32, 64, 128)
    kw.update(ctu=ctu, bit_depth=pick(8, 10, 10), init_qp=int(r.integers(18, 40)), cu_qp_delta=bool(r.integers(0, 2)), max_merge=int(r.integers(1, 7)),
              dual_tree=bool(r.integers(0, 2)), transform_skip=bool(r.integers(0, 2)), deblocking_disabled=r.random() < 0.15,
              beta_offset_div2=int(r.integers(-3, 4)), tc_offset_div2=int(r.integers(-3, 4)), cabac_init_present=bool(r.integers(0, 2)),
              chroma_collocated=(bool(r.integers(0, 2)), bool(r.integers(0, 2))), max_tb64=bool(r.integers(0, 4)))
    kw["bdpcm"] = kw["transform_skip"] and bool(r.integers(0, 2))
    kw["max_gpm"] = int(r.integers(2, kw["max_merge"] + 1)) if kw["max_merge"] >= 2 else 2
    if kw["max_merge"] < 2: kw["gpm"] = False
    kw["max_sub_merge"] = int(r.integers(1 if kw["sbtmvp"] else 0, 6)) if kw["affine"] else 5
    if ctu == 32: kw.update(max_bt_inter=32, max_tt_inter=32)
    if r.random() < 0.3: kw.update(min_cb=8, min_qt_intra=16, min_qt_inter=16, min_qt_intra_c=16)
    mono = r.random() < 0.1
    if mono: kw.update(chroma_format=0, cclm=False, jccr=False, dual_tree=False)
    wc = int(r.integers(4, 10) if big else r.integers(2, 5)); hc = int(r.integers(4, 10) if big else r.integers(2, 5))
    W = wc * ctu - pick(0, 0, 8, 24 if ctu > 32 else 8); H = hc * ctu - pick(0, 0, 8, 16)
    if ctu == 128 and not big: W, H = min(W, 384), min(H, 256)
    kw.update(width=W, height=H)
    rows = -(-H // ctu)
    if rows >= 2 and r.random() < 0.4:
        n = int(r.integers(2, rows + 1)); cut = sorted(r.choice(np.arange(1, rows), size=n - 1, replace=False).tolist())
        sl = [b - a for a, b in zip([0] + cut, cut + [rows])]
        if sl[-1] > sl[-2]: sl[-1], sl[-2] = sl[-2], sl[-1]
        kw.update(slice_rows=tuple(sl), lf_across_slices=bool(r.integers(0, 2)), deblocking_override=bool(r.integers(0, 2)))
    elif r.random() < 0.25 and (wc >= 2 or rows >= 2):
        def cut(n):
            k = int(r.integers(1, min(n, 3) + 1)); c = sorted(r.choice(np.arange(1, n), size=k - 1, replace=False).tolist()) if k > 1 else []
            return tuple(b - a for a, b in zip([0] + c, c + [n]))
        cols, trs = cut(-(-W // ctu)), cut(rows)
        if len(cols) * len(trs) > 1:
            kw.update(tiles=(cols, trs), slice_per_tile=bool(r.integers(0, 2)), lf_across_tiles=bool(r.integers(0, 2)), lf_across_slices=bool(r.integers(0, 2)))
    if not mono and r.random() < 0.4:
        kw.update(chroma_qp_offsets=(int(r.integers(-4, 5)), int(r.integers(-4, 5)), int(r.integers(-4, 5))), slice_chroma_qp_offsets=bool(r.integers(0, 2)))
        if r.random() < 0.5: kw["cu_chroma_qp_offset_list"] = tuple((int(r.integers(-5, 6)), int(r.integers(-5, 6)), int(r.integers(-5, 6))) for _ in range(int(r.integers(1, 5))))
    if r.random() < 0.3: kw["ladf"] = (int(r.integers(-5, 6)), [(int(r.integers(-6, 7)), int(r.integers(20, (1 << kw["bit_depth"]) // 5))) for _ in range(int(r.integers(1, 5)))])
    if not mono and r.random() < 0.4:
        def table():
            start = int(r.integers(-8, 5)); pts = [(int(r.integers(0, 4)), 0) for _ in range(int(r.integers(1, 5)))]
            return (start, [(a, int(r.integers(0, a + 2))) for a, _ in pts])
        kw["chroma_qp_tables"] = tuple(table() for _ in range(pick(1, 3 if kw["jccr"] else 2)))
    kw.update(ph_tool_control=bool(r.integers(0, 2)), parallel_merge_level=int(r.integers(2, 6)), lfnst_scaling_disabled=bool(r.integers(0, 2)))
    if kw["transform_skip"]: kw.update(min_qp_prime_ts=int(r.integers(0, 4)), ts_max_size=int(r.integers(2, 6)))
    if kw.get("chroma_qp_offsets") is not None and r.random() < 0.5: kw["cb_cr_deblock_offsets"] = tuple(int(v) for v in r.integers(-4, 5, size=4))
    structure = pick("gop", "gop", "low_delay", "intra", "gop8")
    if structure == "gop8": kw["dpb_size"] = 8
    pics = gop8(n_gops=pick(1, 2)) if structure == "gop8" else gop4() + (gop4(4, idr=False)[1:] if r.random() < 0.3 else []) if structure == "gop" else low_delay(int(r.integers(3, 7))) if structure == "low_delay" else [vs.Pic(0), vs.Pic(1, idr=True)]
    if r.random() < 0.3 and structure != "intra":
        kw.update(weighted_pred=True, weighted_bipred=bool(r.integers(0, 2)))
        for i, q in enumerate(pics): q["wp"] = seed * 7 + i
    if r.random() < 0.5:
        kw.update(alf=True, ccalf=(not mono) and bool(r.integers(0, 2))); vs.with_alf(pics, r, cc=kw["ccalf"], chroma=not mono)
    if r.random() < 0.4:
        kw["lmcs"] = True; vs.with_lmcs(pics, r, bit_depth=kw["bit_depth"], every=pick(1, 2), chroma=not mono)
        for q in pics:                                                  # (a picture without its own model keeps LMCS off)
            pass
    if kw.get("lmcs") and not kw["max_tb64"] and ctu > 32: kw["ciip"] = False    # (refused by the class: the reference maps residual-free CIIP blocks forward twice there, DecLibReconB200::refuse)
    if r.random() < 0.3:
        kw["scaling_lists"] = True; vs.with_scaling_lists(pics, r, chroma_present=not mono)
    n_slices = len(kw["slice_rows"]) if kw.get("slice_rows") else len(kw["tiles"][0]) * len(kw["tiles"][1]) if kw.get("tiles") and kw.get("slice_per_tile") else 1
    if n_slices > 1 and structure != "intra" and r.random() < 0.5:
        n = n_slices
        for q in pics[1:]: q["slice_types"] = [vs.SLICE_I if (k + q.poc) % 3 == 1 else q.slice_type for k in range(n)]
        if kw.get("weighted_pred"): pass
    for q in pics:
        q["qp"] = kw["init_qp"] + int(r.integers(-4, 5)); q["dep_quant"] = bool(r.integers(0, 2)); q["sign_hiding"] = bool(r.integers(0, 2))
        q["bdof"], q["dmvr"], q["prof"], q["jccr_sign"] = (bool(r.integers(0, 2)) for _ in range(4))
        q["sao"] = (bool(r.integers(0, 4)), bool(r.integers(0, 4))); q["mvd_l1_zero"] = r.random() < 0.2; q["col_from_l0"] = bool(r.integers(0, 2)); q["cabac_init"] = bool(r.integers(0, 2))
    return kw, pics, structure


if __name__ == "__main__":
    first, count = int(sys.argv[1]), int(sys.argv[2])
    oracle = helpers.load_oracle(); bad = 0
    for seed in range(first, first + count):
        kw, pics, structure = random_case(seed, big=len(sys.argv) > 3 and sys.argv[3] == "big")
        tag = f"{seed} {structure} {kw['width']}x{kw['height']} ctu{kw['ctu']} {kw['bit_depth']}b slices={kw.get('slice_rows')} tiles={kw.get('tiles')}{'S' if kw.get('slice_per_tile') else ''} " + "".join(k[0] for k in ("alf", "lmcs", "scaling_lists", "weighted_pred") if kw.get(k))
        try:
            aus, drawn, nb = vs.build_stream(vs.Config(**kw), pics, seed=seed, hash_sei=("md5", "crc", "checksum")[seed % 3] if seed & 1 else None)
        except (vs.DecodeError, AssertionError) as e:
            print(tag, "not drawn:", str(e)[-220:].replace("\n", " "), flush=True); continue
        stock = vs.decode(vs.REF_SO, aus, frame_samples=kw['width'] * kw['height'] * 2)
        d0 = _diff(drawn, stock)
        try:
            sw, log = su.decode_swapped_cpu(aus, oracle, threads=int(seed % 3 == 0) * 3 + 1, async_finish=bool(seed & 1), frame_samples=kw['width'] * kw['height'] * 2)
            d1 = _diff(sw, stock)
        except Exception as e:
            d1 = "FAILED " + str(e)[-300:].replace("\n", " ")
        ok = d0 == [0] * len(aus) and d1 == [0] * len(aus)
        bad += not ok
        print(tag, "bins", sum(nb), "OK" if ok else f"DIFF writer {d0} seam {d1}", flush=True)
    sys.exit(1 if bad else 0)
