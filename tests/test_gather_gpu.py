"""Multi-GPU display-order gather over NCCL: two ranks, each with its own device context, hand finished frames (DPB slots) to
vvdec_b200.gather.FrameGather through b200_get_frame_device_async on a side stream; rank 0 must hold every frame of every GOP in display order,
bit-exact.  Needs two GPUs (skipped on a one-GPU box; the same class runs on gloo in tests/test_gop_shard_cpu.py)."""
import os, sys, subprocess, textwrap
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_two_rank_nccl_gather_display_order(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "g.py"
    script.write_text(textwrap.dedent(f"""
        import sys, os, ctypes as C; sys.path.insert(0, {ROOT!r})
        import numpy as np, torch, torch.distributed as dist
        import vvdec_b200
        from vvdec_b200 import abi, gather
        local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        r, w = dist.get_rank(), dist.get_world_size()
        lib = vvdec_b200.lib()
        W, H = 416, 240
        g = abi.make_geom(W, H, 10)
        ctx = C.c_void_p(); vvdec_b200.check(lib.b200_ctx_create(C.byref(ctx), C.byref(g), 4, 2, local))
        lengths = [3, 2, 4, 1]                                   # GOPs 0, 2 on rank 0; 1, 3 on rank 1
        numel = W * H * 3                                         # bytes of a 16-bit 4:2:0 frame
        G = gather.FrameGather(r, w, lengths, numel, torch.device("cuda", local))
        side = torch.cuda.Stream()
        def frame(k, i):                                          # the content of frame i of GOP k
            rng = np.random.default_rng(100 * k + i)
            return [rng.integers(0, 1024, size=(H, W), dtype=np.int16), rng.integers(0, 1024, size=(H // 2, W // 2), dtype=np.int16), rng.integers(0, 1024, size=(H // 2, W // 2), dtype=np.int16)]
        li = 0
        for k in G.assignment[r]:
            for i in range(lengths[k]):
                fr = frame(k, i)                                  # kept alive across the call: plane_ptrs holds raw pointers
                vvdec_b200.check(lib.b200_ctx_load_slot(ctx, li % 4, abi.plane_ptrs(fr)))
                base = G.slot(li).data_ptr()
                pl = (C.c_void_p * 3)(base, base + 2 * W * H, base + 2 * (W * H + W * H // 4))
                vvdec_b200.check(lib.b200_get_frame_device_async(ctx, li % 4, pl, C.c_void_p(side.cuda_stream)))
                with torch.cuda.stream(side): G.push(li)
                side.synchronize()                                # the slot is reused below: this test has 4 slots for up to 5 frames
                li += 1
        with torch.cuda.stream(side): store = G.finish()
        torch.cuda.synchronize()
        if r == 0:
            got = store.cpu().numpy(); d = 0
            for k in range(len(lengths)):
                for i in range(lengths[k]):
                    want = np.concatenate([p.ravel() for p in frame(k, i)]).view(np.uint8)
                    assert np.array_equal(got[d], want), (k, i)
                    d += 1
            print("GATHER_OK", d)
        dist.barrier(); lib.b200_ctx_destroy(ctx); dist.destroy_process_group()
    """))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29621", str(script)], capture_output=True, text=True, timeout=600)
    if os.environ.get("B200_TEST_LOG"): open(os.environ["B200_TEST_LOG"], "w").write(out.stdout + "\n---- stderr\n" + out.stderr)
    assert "GATHER_OK 10" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
