"""GOP sharding for multi-GPU decoding: closed GOPs (IRAP-to-IRAP coded video sequences) are independent,
so GOP k goes to rank k mod N with no exchange during decode.  Host-side logic only."""
from typing import List, Sequence

IDR_NAL_TYPES = {7, 8}       # IDR_W_RADL, IDR_N_LP (vvdecNalType, include/vvdec/vvdec.h.in)
CRA_NAL_TYPE, RASL_NAL_TYPE = 9, 3
IRAP_NAL_TYPES = IDR_NAL_TYPES | {CRA_NAL_TYPE}


def split_gops(nal_types: Sequence[int]) -> List[range]:
    """Cut an access-unit sequence (one NAL unit type per AU) where a decoder can start without anything decoded before: at IDR pictures, and at a
    CRA picture only if no RASL pictures follow it (RASL pictures reference pictures that precede the CRA: an open GOP stays with its predecessor,
    a rank that started at the CRA would have to drop them — DecLibParser.cpp:1621 isRandomAccessSkipPicture — and the display order would have holes).
    Returns AU index ranges, one per independently decodable GOP."""
    starts = []
    for i, t in enumerate(nal_types):
        if t in IDR_NAL_TYPES: starts.append(i)
        elif t == CRA_NAL_TYPE:
            j = i + 1
            while j < len(nal_types) and nal_types[j] not in IRAP_NAL_TYPES and nal_types[j] != RASL_NAL_TYPE: j += 1
            if j >= len(nal_types) or nal_types[j] != RASL_NAL_TYPE: starts.append(i)
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
