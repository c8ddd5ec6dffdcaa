"""GOP sharding for multi-GPU decoding: closed GOPs (IRAP-to-IRAP coded video sequences) are independent,
so GOP k goes to rank k mod N with no exchange during decode.  Host-side logic only."""
from typing import List, Sequence

IRAP_NAL_TYPES = {7, 8, 9}   # IDR_W_RADL, IDR_N_LP, CRA (vvdecNalType, include/vvdec/vvdec.h.in)


def split_gops(nal_types: Sequence[int]) -> List[range]:
    """Cut an access-unit sequence (one NAL unit type per AU) at IRAP pictures. Returns AU index ranges, one per GOP."""
    starts = [i for i, t in enumerate(nal_types) if t in IRAP_NAL_TYPES]
    if not starts or starts[0] != 0:
        starts = [0] + starts
    return [range(a, b) for a, b in zip(starts, starts[1:] + [len(nal_types)]) if b > a]


def assign(num_gops: int, world: int) -> List[List[int]]:
    """GOP k -> rank k mod world."""
    return [[k for k in range(num_gops) if k % world == r] for r in range(world)]


def output_order(assignment: List[List[int]], gop_lengths: Sequence[int]):
    """(rank, local frame index) for every frame in display order — what the final gather follows."""
    pos = {}
    for r, gops in enumerate(assignment):
        off = 0
        for k in gops:
            pos[k] = (r, off); off += gop_lengths[k]
    out = []
    for k in range(len(gop_lengths)):
        r, off = pos[k]
        out += [(r, off + i) for i in range(gop_lengths[k])]
    return out
