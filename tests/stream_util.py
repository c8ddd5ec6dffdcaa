"""Bitstream-level test plumbing: a stream written by oracle/vvc_stream.py decoded by (a) the stock reference and (b) the reference with the drop-in class behind
the DecLibRecon seam (oracle/_ref/libvvdec_swapped.so).  On a machine without a GPU (b) runs its host stages for real and the oracle chain in place of the device
(DecLibReconB200::TestHooks); on the GPU box it is the product path."""
import ctypes as C, numpy as np
from tests import helpers
from vvdec_b200 import abi
from oracle import vvc_stream as vs

NUM_SLOTS = 17                                                   # DecLibReconB200::m_dpbSlots

_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(abi.Picture), C.POINTER(abi.Geom), C.POINTER(C.c_int32), C.c_size_t,
                    C.POINTER(C.POINTER(C.c_int16)), C.POINTER(C.c_ssize_t), C.c_int)


_LOAD = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.POINTER(C.c_int16)), C.POINTER(C.c_ssize_t), C.POINTER(abi.Geom))


def swapped_lib():
    lib = vs._lib(vs.SWAP_SO)
    lib.swapped_set_hooks.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.swapped_set_async_finish.argtypes = [C.c_int]
    return lib


class OracleDevice:
    """Stands where the device would: reconstructs every flattened picture with the oracle chain and keeps the decoded-picture buffer by slot."""
    def __init__(self, oracle):
        self.oracle, self.dpb, self.log, self.error = oracle, None, [], None
        self.keep, self.pics = False, {}                            # keep: the flattened lists of every picture by POC (diagnosis)
        self.cb = _HOOK(self._picture); self.load_cb = _LOAD(self._load)

    def _slots(self, g):
        if self.dpb is None or self.dpb[0][0].shape != (g.height, g.width):          # (first picture, or the context was rebuilt for another geometry)
            self.dpb = [[np.zeros((g.height, g.width), np.int16), np.zeros((g.height // 2, g.width // 2), np.int16), np.zeros((g.height // 2, g.width // 2), np.int16)]
                        for _ in range(NUM_SLOTS)]

    def _load(self, user, slot, planes, strides, geom):
        """b200_ctx_load_slot_strided: a reference the device does not hold is taken from its host planes"""
        try:
            g = geom.contents; self._slots(g)
            for c in range(3 if g.chromaFormat else 1):
                h, w = self.dpb[slot][c].shape
                src = np.ctypeslib.as_array(planes[c], shape=((h - 1) * strides[c] + w,))
                self.dpb[slot][c] = np.lib.stride_tricks.as_strided(src, shape=(h, w), strides=(strides[c] * 2, 2)).copy()
        except BaseException:
            import traceback; self.error = traceback.format_exc()

    def _picture(self, user, lists, geom, dmvr, ndmvr, planes, strides, poc):
        try:
            g = geom.contents; st = lists.contents
            self._slots(g)
            pic = helpers.picture_from_struct(st, g, None)
            out, dm = helpers.oracle_decompress(self.oracle, g, self.dpb, pic)
            self.dpb[st.dstSlot] = out
            if self.keep: self.pics[poc] = pic
            n = min(int(ndmvr), len(dm))
            for i in range(n): dmvr[2 * i], dmvr[2 * i + 1] = int(dm[i][0]), int(dm[i][1])
            for c in range(3 if g.chromaFormat else 1):
                h, w = out[c].shape
                dst = np.ctypeslib.as_array(planes[c], shape=((h - 1) * strides[c] + w,))
                np.lib.stride_tricks.as_strided(dst, shape=(h, w), strides=(strides[c] * 2, 2))[...] = out[c]
            self.log.append(dict(poc=poc, slot=int(st.dstSlot), pus=int(st.numPus), tus=int(st.numTus), intra=int(st.numIntraTus), flags=int(st.flags), scaling=int(st.numScaling), wp=int(st.numWp), lfSlices=int(st.numLfSlices)))
        except BaseException as e:                                # never unwind through the C++ frames
            import traceback; self.error = traceback.format_exc()


def decode_swapped_cpu(aus, oracle, threads=1, keep=None, async_finish=False, **kw):
    """The stream through the swapped build without a device: glue host stages + oracle chain.  Returns (frames, per-picture log).
    keep: a dict that receives the flattened work lists of every picture by POC."""
    lib = swapped_lib(); dev = OracleDevice(oracle)
    if keep is not None: dev.keep, dev.pics = True, keep
    lib.swapped_set_hooks(1, C.cast(dev.cb, C.c_void_p), C.cast(dev.load_cb, C.c_void_p), None)
    lib.swapped_set_async_finish(int(async_finish))            # pictures complete in a pool task (DecLibReconB200::setAsyncFinish) instead of in waitForPrevDecompressedPic()
    try:
        frames = vs.decode(vs.SWAP_SO, aus, threads=threads, **kw)
    finally:
        lib.swapped_set_hooks(1, None, None, None); lib.swapped_set_async_finish(0)
    assert dev.error is None, dev.error
    return frames, dev.log


def decode_swapped_device(aus, threads=8, async_finish=False, **kw):
    """The stream through the swapped build on the product path (GPU)."""
    lib = swapped_lib(); lib.swapped_set_hooks(0, None, None, None); lib.swapped_set_async_finish(int(async_finish))
    try: return vs.decode(vs.SWAP_SO, aus, threads=threads, **kw)
    finally: lib.swapped_set_async_finish(0)


def corruption_run(seed, n, threads):
    """n streams with 1-3 flipped bits in a later access unit through the swapped build (oracle device); prints `ok` per stream that came back — with an error code
    or with frames — and exits.  Run in a child process (tests/test_stream_cpu.py): a crash or a hang of the class's error paths must not take the test run down."""
    import os
    from tests.test_stream_cpu import ALL, gop4, low_delay
    oracle = helpers.load_oracle(); rng = np.random.default_rng(seed)
    for it in range(n):
        aus, _, _ = vs.build_stream(vs.Config(**dict(ALL, sao=True)), low_delay(5) if it % 2 else gop4(), seed=int(rng.integers(1, 1 << 20)))
        k = int(rng.integers(1, len(aus))); au = bytearray(aus[k])
        for _ in range(int(rng.integers(1, 4))):
            au[int(rng.integers(12, len(au)))] ^= 1 << int(rng.integers(0, 8))
        bad = list(aus); bad[k] = bytes(au)
        try:
            frames, _ = decode_swapped_cpu(bad, oracle, threads=threads); print("ok frames", len(frames), flush=True)
        except vs.DecodeError as e:
            print("ok error", str(e).split("|")[0].strip(), flush=True)
    os._exit(0)


def device_child(in_path, out_path, threads, async_finish):
    """child process of decode_swapped_device_guarded(): decodes on the device path and stores the frames"""
    import os, pickle
    aus, kw = pickle.load(open(in_path, "rb"))
    try:
        frames = decode_swapped_device(aus, threads=threads, async_finish=async_finish, **kw)
        pickle.dump(dict(frames=frames, hash_errors=vs.decode.hash_errors), open(out_path, "wb"))
    except vs.DecodeError as e:
        pickle.dump(dict(error=str(e)), open(out_path, "wb"))
    os._exit(0)


def decode_swapped_device_guarded(aus, threads=4, async_finish=False, timeout=240, **kw):
    """decode_swapped_device() in a child process with a time limit: a crash or a hang on the device path fails the one test instead of taking the test run down.
    Returns (frames, hash errors)."""
    import os, pickle, subprocess, sys, tempfile
    with tempfile.TemporaryDirectory() as d:
        pickle.dump((aus, kw), open(os.path.join(d, "in.pkl"), "wb"))
        p = subprocess.run([sys.executable, "-m", "tests.stream_util", "device", os.path.join(d, "in.pkl"), os.path.join(d, "out.pkl"), str(threads), str(int(async_finish))],
                           capture_output=True, text=True, timeout=timeout, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert p.returncode == 0 and os.path.exists(os.path.join(d, "out.pkl")), f"device decode died (rc {p.returncode}): {p.stderr[-600:]}"
        r = pickle.load(open(os.path.join(d, "out.pkl"), "rb"))
    assert "error" not in r, r.get("error")
    return r["frames"], r["hash_errors"]


if __name__ == "__main__":
    import sys
    if sys.argv[1] == "device": device_child(sys.argv[2], sys.argv[3], int(sys.argv[4]), bool(int(sys.argv[5])))
    corruption_run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))
