"""Host-side multi-GPU logic on CPU: GOP cutting / assignment, and a world_size-2 gloo run in which each rank 'decodes' its GOPs
(frame checksums stand in for frames) and rank 0 gathers them in display order."""
import os, sys, subprocess, textwrap
import numpy as np
from vvdec_b200 import gop_shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_and_assign():
    nal = [8, 0, 0, 0, 9, 0, 0, 7, 0, 0, 0, 0]
    gops = gop_shard.split_gops(nal)
    assert [list(g) for g in gops] == [[0, 1, 2, 3], [4, 5, 6], [7, 8, 9, 10, 11]]
    # an open GOP (CRA followed by RASL pictures) is not a cut point; a CRA without leading pictures is
    assert [list(g) for g in gop_shard.split_gops([8, 0, 0, 9, 3, 3, 0, 9, 0, 0])] == [[0, 1, 2, 3, 4, 5, 6], [7, 8, 9]]
    a = gop_shard.assign(5, 2)
    assert a == [[0, 2, 4], [1, 3]]
    order = gop_shard.output_order(gop_shard.assign(3, 2), [4, 3, 5])
    assert order[:4] == [(0, 0), (0, 1), (0, 2), (0, 3)] and order[4] == (1, 0) and order[7] == (0, 4) and len(order) == 12


def test_gloo_two_ranks_gather(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import sys; sys.path.insert(0, {ROOT!r})
        import torch, torch.distributed as dist
        from vvdec_b200 import gop_shard
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        lengths = [4, 3, 5, 2]
        mine = gop_shard.assign(len(lengths), w)[r]
        frames = torch.tensor([1000 * k + i for k in mine for i in range(lengths[k])], dtype=torch.int64)   # frame "checksums"
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(w)]
        dist.all_gather(sizes, torch.tensor([len(frames)]))
        pad = torch.zeros(int(max(s.item() for s in sizes)), dtype=torch.int64); pad[:len(frames)] = frames
        allf = [torch.zeros_like(pad) for _ in range(w)]
        dist.all_gather(allf, pad)
        if r == 0:
            order = gop_shard.output_order(gop_shard.assign(len(lengths), w), lengths)
            got = [int(allf[rr][i]) for rr, i in order]
            want = [1000 * k + i for k in range(len(lengths)) for i in range(lengths[k])]
            assert got == want, (got, want)
            print("OK")
        dist.destroy_process_group()
    """))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29611", str(script)], capture_output=True, text=True, timeout=300)
    assert "OK" in out.stdout, out.stdout + out.stderr


def test_gloo_frame_gather_display_order(tmp_path):
    """vvdec_b200.gather.FrameGather (the class bench.py uses over NCCL) on gloo, 3 ranks: frames arrive on rank 0 in display order."""
    script = tmp_path / "g.py"
    script.write_text(textwrap.dedent(f"""
        import sys; sys.path.insert(0, {ROOT!r})
        import torch, torch.distributed as dist
        from vvdec_b200 import gather, gop_shard
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        lengths = [4, 3, 5, 2, 3]
        G = gather.FrameGather(r, w, lengths, 8, "cpu", dtype=torch.int32)
        li = 0
        for k in G.assignment[r]:
            for i in range(lengths[k]):
                G.slot(li).fill_(1000 * k + i); G.push(li); li += 1
        store = G.finish()
        if r == 0:
            want = [1000 * k + i for k in range(len(lengths)) for i in range(lengths[k])]
            assert [int(v) for v in store[:, 0]] == want and bool((store == store[:, :1]).all()), store[:, 0]
            print("OK")
        dist.barrier(); dist.destroy_process_group()
    """))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3", "--master-addr", "127.0.0.1",
                          "--master-port", "29613", str(script)], capture_output=True, text=True, timeout=300)
    assert "OK" in out.stdout, out.stdout + out.stderr
