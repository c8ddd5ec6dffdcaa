"""Display-order gather of finished frames to rank 0 (BASELINE north_star: "NCCL only for the final gather").

GOPs are decoded on different GPUs (gop_shard.assign); the consumer of the pictures (a display, an encoder, a file writer) sits on rank 0 and wants them
in display order.  Every other rank sends each finished frame to rank 0 as soon as it is final; rank 0 has the receives posted in display order per
peer.  One two-rank process group per peer keeps the peers' transfers independent (a frame of rank 2 never waits behind rank 1's queue), and the sends
run on a side stream, so the transfers overlap the decoding of the following pictures.  Backend-agnostic: NCCL over NVLink on the GPUs, gloo in the
CPU tests."""
from typing import List, Sequence
import torch, torch.distributed as dist
from . import gop_shard


class FrameGather:
    def __init__(self, rank: int, world: int, gop_lengths: Sequence[int], frame_numel: int, device, dtype=torch.uint8):
        # frames travel as bytes: torch's NCCL process group has no 16-bit integer type
        self.rank, self.world, self.numel = rank, world, frame_numel
        self.assignment = gop_shard.assign(len(gop_lengths), world)
        self.order = gop_shard.output_order(self.assignment, gop_lengths)              # display index -> (rank, local frame index)
        self.display_of = {rl: d for d, rl in enumerate(self.order)}
        self.local_frames = sum(gop_lengths[k] for k in self.assignment[rank])
        self.groups = {}
        if world > 1:
            for r in range(1, world):                                                  # every rank creates every group (collective call)
                g = dist.new_group([0, r])
                if rank in (0, r): self.groups[r] = g
        self.reqs: List = []
        self.store = None
        if rank == 0:
            self.store = torch.empty((len(self.order), frame_numel), dtype=dtype, device=device)
            for r in range(1, world):                                                  # receives of one peer in that peer's sending order
                for d, (rr, _) in enumerate(self.order):
                    if rr == r: self.reqs.append(dist.irecv(self.store[d], src=r, group=self.groups[r]))
        else:
            self.staging = torch.empty((self.local_frames, frame_numel), dtype=dtype, device=device)   # a send buffer per frame: nothing waits for a buffer

    def slot(self, local_index: int) -> torch.Tensor:
        """Where the frame with this local (decoding-order) index has to be written: the display store on rank 0, a send buffer elsewhere."""
        if self.rank == 0: return self.store[self.display_of[(0, local_index)]]
        return self.staging[local_index]

    def push(self, local_index: int):
        """The frame in slot(local_index) is complete (on the current stream): send it."""
        if self.rank != 0: self.reqs.append(dist.isend(self.staging[local_index], dst=0, group=self.groups[self.rank]))

    def finish(self):
        for r in self.reqs: r.wait()
        self.reqs = []
        return self.store

    def bytes_received(self, itemsize=1) -> int:
        return sum(1 for r, _ in self.order if r != 0) * self.numel * itemsize if self.rank == 0 else 0
