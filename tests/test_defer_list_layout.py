"""Index arithmetic of the in-place hand-over between bulk2_pair_kernel and bulk2_second_kernel (csrc/bulk2.cuh): warp
`part` of `parts` walks a segment of the deferred list in 32-entry batches part, part + parts, ...; the m-th entry it
leaves is written where the m-th entry of its walk was, and the second pass reads `left_n` entries from the same places.
Pure host arithmetic (the kernels themselves are covered by the -m gpu parity tests)."""
import random

import pytest


def walk(part, parts, cnt):
    """(turn, lane, index) of every entry warp `part` reads from a segment of `cnt` entries (bulk2_pair_kernel)"""
    out, turn, i0 = [], 0, part * 32
    while i0 < cnt:
        out += [(turn, lane, i0 + lane) for lane in range(32) if i0 + lane < cnt]
        i0 += parts * 32
        turn += 1
    return out


def write_pos(part, parts, m):
    return part * 32 + (m >> 5) * parts * 32 + (m & 31)          # list[part * 32 + (m >> 5) * parts * 32 + (m & 31)] = rec


def second_pass_reads(part, parts, left_n):
    """indices bulk2_second_kernel reads for (segment, part) when part_count says left_n"""
    cnt = part * 32 + ((left_n + 31) // 32) * parts * 32
    out, turn, i0 = [], 0, part * 32
    while i0 < cnt:
        out += [i0 + lane for lane in range(32) if turn * 32 + lane < left_n]
        i0 += parts * 32
        turn += 1
    return out


@pytest.mark.parametrize("parts", [1, 2, 3, 4, 7, 16])
@pytest.mark.parametrize("cnt", [0, 1, 31, 32, 33, 1000, 4097])
def test_in_place_hand_over(parts, cnt):
    rng = random.Random(parts * 100003 + cnt)
    written = {}
    for part in range(parts):
        entries = walk(part, parts, cnt)
        m = 0
        for turn, lane, idx in entries:
            if rng.random() < 0.3:                               # this entry is left for the second pass
                pos = write_pos(part, parts, m)
                # the place belongs to this warp's own walk and to a batch it has read already (this turn's or earlier)
                assert pos % (parts * 32) // 32 == part and pos // (parts * 32) <= turn
                assert pos < max(cnt, 1) + 0 or pos < part * 32 + (turn + 1) * parts * 32
                assert pos not in written
                written[pos] = (part, m)
                m += 1
        assert second_pass_reads(part, parts, m) == [write_pos(part, parts, k) for k in range(m)]
    # every warp's walk covers the segment exactly once between them
    seen = sorted(idx for part in range(parts) for _, _, idx in walk(part, parts, cnt))
    assert seen == list(range(cnt))
