"""Host-side contract of the peer exchange + find_device_split (no GPU): the slot a rank's slice takes in a peer's
buffer (csrc/api.cu yume_pdq_cuda_exchange_post_bytes) and the needle-index map of the packed batch (csrc/common.cuh
NeedleMap) together reproduce the needle numbering of an all-gather, for every world size the mailboxes allow."""
import re
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def slot_of(rank: int, peer: int) -> int:
    """where rank's slice lands in peer's buffer (api.cu: `ex->rank < p ? ex->rank : ex->rank - 1`)"""
    return rank if rank < peer else rank - 1


def needle_map(j: int, base: int, split_at: int, gap: int) -> int:
    """common.cuh NeedleMap::operator()"""
    return base + j + (gap if j >= split_at else 0)


@pytest.mark.parametrize("world", range(2, 17))
@pytest.mark.parametrize("k8", [8, 1024])
def test_packed_batch_reports_all_gather_indices(world, k8):
    base = 5 * world * k8                        # a chunk offset, as ShardedMatcher.screen uses
    for me in range(world):
        seen = {}
        # own needles: needle_base = base + me * k8
        for i in range(k8):
            seen[needle_map(i, base + me * k8, 0xFFFFFFFF, 0)] = (me, i)
        # the packed batch of the others: slot s holds rank (s if s < me else s + 1)
        for r in range(world):
            if r == me:
                continue
            s = slot_of(r, me)
            for i in (0, 1, k8 - 1):
                j = s * k8 + i
                g = needle_map(j, base, me * k8, k8)
                assert g == base + r * k8 + i, (world, me, r, i)
                seen[g] = (r, i)
        assert min(seen) == base and max(seen) == base + world * k8 - 1
    # every slot of a peer's buffer is taken by exactly one rank
    for p in range(world):
        assert sorted(slot_of(r, p) for r in range(world) if r != p) == list(range(world - 1))


def test_sources_still_say_so():
    """the two formulas above are the ones in the sources"""
    api = open(os.path.join(ROOT, "yume_pdq_b200", "csrc", "api.cu")).read()
    assert re.search(r"ex->rank < p \? ex->rank : ex->rank - 1", api)
    common = open(os.path.join(ROOT, "yume_pdq_b200", "csrc", "common.cuh")).read()
    assert re.search(r"return base \+ j \+ \(j >= split_at \? gap : 0u\);", common)
