"""Device-side output conversion (b200_get_frame_fmt_async) against the oracle's restatement of vvdecapp's writers."""
import ctypes as C
import numpy as np
import pytest
import vvdec_b200
from vvdec_b200 import abi, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,H", [(416, 240), (24, 16), (1920, 1080)])
def test_frame_formats(b200, oracle, W, H):
    rng = np.random.default_rng(W)
    bd = 10
    g = abi.make_geom(W, H, bd)
    ctx = C.c_void_p()
    vvdec_b200.check(b200.b200_ctx_create(C.byref(ctx), C.byref(g), 2, 1, -1))
    try:
        pl = synth.noise_planes(rng, W, H, bd)
        pl[0][0, :4] = [0, 1023, 1, 1022]
        vvdec_b200.check(b200.b200_ctx_load_slot(ctx, 1, abi.plane_ptrs(pl)))
        for fmt in (1, 2, 0):
            outs = [np.zeros(b200.b200_frame_bytes(C.byref(g), fmt, c), np.uint8) for c in range(3)]
            ptrs = (C.c_void_p * 3)(*[o.ctypes.data for o in outs])
            t = b200.b200_get_frame_fmt_async(ctx, 1, fmt, ptrs); assert t >= 0, b200.b200_last_error()
            vvdec_b200.check(b200.b200_frame_wait(ctx, t))
            for c in range(3):
                w, h = (W, H) if c == 0 else (W // 2, H // 2)
                if fmt == 0: want = pl[c].view(np.uint8).reshape(-1)
                else:
                    want = np.zeros(len(outs[c]), np.uint8)
                    if fmt == 1: oracle.orc_pack_pyuv(pl[c], pl[c].shape[1], w, h, want)
                    else: oracle.orc_narrow8(pl[c], pl[c].shape[1], w, h, bd, want)
                assert np.array_equal(outs[c], want), (fmt, c)
    finally:
        b200.b200_ctx_destroy(ctx)


@pytest.mark.parametrize("W,H,bd", [(416, 240, 10), (24, 16, 10), (1920, 1080, 10), (416, 240, 8), (3840, 2160, 10), (72, 40, 12)])
def test_frame_hash(b200, oracle, W, H, bd):
    """b200_frame_hash_async (CRC by chunk reduction + GF(2) polynomial recombination, checksum by atomic sums) against the oracle's
    serial restatement of calcCRC / calcChecksum, which tests/test_output_oracle_vs_ref.py pins to the reference."""
    rng = np.random.default_rng(W + bd)
    g = abi.make_geom(W, H, bd)
    ctx = C.c_void_p()
    vvdec_b200.check(b200.b200_ctx_create(C.byref(ctx), C.byref(g), 2, 1, -1))
    try:
        pl = synth.noise_planes(rng, W, H, bd)
        pl[0][0, :4] = [0, (1 << bd) - 1, 255 % (1 << bd), 256 % (1 << bd)]
        vvdec_b200.check(b200.b200_ctx_load_slot(ctx, 1, abi.plane_ptrs(pl)))
        tickets = []
        for method, n in ((1, 2), (2, 4)):
            dig = np.full(12, 0xEE, np.uint8)
            t = b200.b200_frame_hash_async(ctx, 1, method, dig.ctypes.data); assert t >= 0, b200.b200_last_error()
            tickets.append((t, method, n, dig))
        for t, method, n, dig in tickets:
            vvdec_b200.check(b200.b200_frame_wait(ctx, t))
            for c in range(3):
                want = np.zeros(4, np.uint8)
                assert oracle.orc_plane_hash(method, bd, pl[c], pl[c].shape[1], W >> (c > 0), H >> (c > 0), want) == n
                assert np.array_equal(dig[c * n:(c + 1) * n], want[:n]), (method, c)
        assert b200.b200_frame_hash_async(ctx, 1, 0, dig.ctypes.data) == -4 and b"MD5" in b200.b200_last_error()
    finally:
        b200.b200_ctx_destroy(ctx)


def test_frame_to_device_buffer_on_a_side_stream(b200):
    """b200_get_frame_device_async: a DPB slot copied into caller-owned device memory on the caller's stream (what the multi-GPU gather sends from)."""
    import torch
    W, H = 416, 240
    g = abi.make_geom(W, H, 10)
    ctx = C.c_void_p(); vvdec_b200.check(b200.b200_ctx_create(C.byref(ctx), C.byref(g), 4, 2, 0))
    rng = np.random.default_rng(5)
    side = torch.cuda.Stream()
    for it in range(3):
        planes = synth.noise_planes(rng, W, H, 10)
        vvdec_b200.check(b200.b200_ctx_load_slot(ctx, it % 4, abi.plane_ptrs(planes)))
        dst = torch.zeros(W * H * 3, dtype=torch.uint8, device="cuda")
        base = dst.data_ptr()
        pl = (C.c_void_p * 3)(base, base + 2 * W * H, base + 2 * (W * H + W * H // 4))
        vvdec_b200.check(b200.b200_get_frame_device_async(ctx, it % 4, pl, C.c_void_p(side.cuda_stream)))
        side.synchronize()
        want = np.concatenate([p.ravel() for p in planes]).view(np.uint8)
        assert np.array_equal(dst.cpu().numpy(), want), it
    b200.b200_ctx_destroy(ctx)
