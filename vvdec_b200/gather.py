"""Display-order gather of finished frames to rank 0 (BASELINE north_star: "NCCL only for the final gather").

GOPs are decoded on different GPUs (gop_shard.assign); the consumer of the pictures (a display, an encoder, a file writer) sits on rank 0 and wants them
in display order.  Every other rank sends each finished frame to rank 0 as soon as it is final, rank 0 receives it straight into the frame's place in the
display-order store.  Backend-agnostic: NCCL over NVLink on the GPUs, gloo in the CPU tests."""
from typing import List, Sequence
import torch, torch.distributed as dist
from . import gop_shard


class FrameGather:
    """Per-step exchange on the world communicator: when a rank has finished its i-th local frame it calls push(i); rank 0's push(i) posts the receives of
    every peer's i-th frame, a peer's push(i) posts the send of its own — one batched P2P group per call (ncclGroupStart/End under NCCL), issued in the same
    order on all ranks, so the matching operations line up without any handshake.  The transfers run on the communicator's stream behind the stream that
    called push() (the side stream the frame was copied on) and overlap the reconstruction of the following pictures; finish() posts what is left (peers
    with more frames than rank 0) and waits."""
    def __init__(self, rank: int, world: int, gop_lengths: Sequence[int], frame_numel: int, device, dtype=torch.uint8, display_of=None):
        # frames travel as bytes: torch's NCCL process group has no 16-bit integer type
        self.rank, self.world, self.numel = rank, world, frame_numel
        if display_of is None:
            self.assignment = gop_shard.assign(len(gop_lengths), world)
            self.order = gop_shard.output_order(self.assignment, gop_lengths)          # display index -> (rank, local frame index)
            self.display_of = {rl: d for d, rl in enumerate(self.order)}
            self.frames_of = [sum(gop_lengths[k] for k in self.assignment[r]) for r in range(world)]
        else:
            # explicit map (rank, n-th frame that rank pushes) -> display index, the same on every rank: for ranks that reconstruct several GOPs side by side
            # (GOP-parallel lanes), whose frames do not finish in GOP order
            self.display_of = dict(display_of)
            self.order = [rl for rl, _ in sorted(self.display_of.items(), key=lambda kv: kv[1])]
            self.frames_of = [sum(1 for (r, _) in self.display_of if r == q) for q in range(world)]
        self.local_frames = self.frames_of[rank]
        self.reqs: List = []
        self.posted = 0                                                               # rank 0: local indices whose receives are posted
        self.store = None
        if rank == 0: self.store = torch.empty((len(self.order), frame_numel), dtype=dtype, device=device)
        else: self.staging = torch.empty((self.local_frames, frame_numel), dtype=dtype, device=device)   # a send buffer per frame: nothing waits for a buffer

    def slot(self, local_index: int) -> torch.Tensor:
        """Where the frame with this local (decoding-order) index has to be written: the display store on rank 0, a send buffer elsewhere."""
        if self.rank == 0: return self.store[self.display_of[(0, local_index)]]
        return self.staging[local_index]

    def _post(self, i: int):
        if self.world == 1: return
        if self.rank == 0: ops = [dist.P2POp(dist.irecv, self.store[self.display_of[(r, i)]], r) for r in range(1, self.world) if i < self.frames_of[r]]
        else: ops = [dist.P2POp(dist.isend, self.staging[i], 0)]
        if ops: self.reqs.extend(dist.batch_isend_irecv(ops))

    def push(self, local_index: int):
        """The frame in slot(local_index) is complete (on the current stream): exchange the frames with this index."""
        if self.rank == 0:
            while self.posted <= local_index: self._post(self.posted); self.posted += 1
        else: self._post(local_index)

    def finish(self):
        if self.rank == 0:
            while self.posted < max(self.frames_of): self._post(self.posted); self.posted += 1
        for r in self.reqs: r.wait()
        self.reqs = []
        return self.store

    def bytes_received(self, itemsize=1) -> int:
        return sum(1 for r, _ in self.order if r != 0) * self.numel * itemsize if self.rank == 0 else 0
