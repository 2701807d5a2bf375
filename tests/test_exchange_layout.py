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


# ---- the "never two exchanges ahead" argument of yume_pdq_cuda.h / ShardedMatcher.screen, checked on a model ----------
def _simulate(world: int, steps: int, seed: int):
    """Every rank runs, in stream order:  post(c) -> own_match(c) -> wait(c) -> others_begin(c) .. others_end(c)  for
    c = 1 .. steps, with two buffers alternating by c & 1 (csrc/api.cu).  A random scheduler interleaves the ranks; a push
    is split into one event per peer (the kernel's stores reach the peers at different times) followed by the flag
    stores.  Returns the number of pushes that landed in a buffer a peer was still reading (must be 0)."""
    import random
    rnd = random.Random(seed)
    flags = [[0] * world for _ in range(world)]            # flags[p][r]: last sequence number rank r pushed to p
    reading = [None] * world                               # parity of the buffer rank p is matching from, if any
    violations = 0
    progs = []
    for r in range(world):
        ops = []
        for c in range(1, steps + 1):
            peers = [p for p in range(world) if p != r]
            rnd.shuffle(peers)
            ops += [("data", c, p) for p in peers] + [("flag", c, p) for p in peers]
            ops += [("own", c, None), ("wait", c, None), ("begin", c, None), ("end", c, None)]
        progs.append(ops)
    pc = [0] * world
    while any(pc[r] < len(progs[r]) for r in range(world)):
        runnable = []
        for r in range(world):
            if pc[r] == len(progs[r]):
                continue
            kind, c, p = progs[r][pc[r]]
            if kind == "wait" and any(flags[r][q] < c for q in range(world) if q != r):
                continue
            runnable.append(r)
        assert runnable, "deadlock"
        r = rnd.choice(runnable)
        kind, c, p = progs[r][pc[r]]
        pc[r] += 1
        if kind == "data":
            if reading[p] == (c & 1):
                violations += 1
        elif kind == "flag":
            flags[p][r] = c
        elif kind == "begin":
            reading[r] = c & 1
        elif kind == "end":
            reading[r] = None
    return violations


@pytest.mark.parametrize("world", [2, 3, 8])
def test_two_buffers_suffice_without_a_per_step_collective(world):
    for seed in range(40):
        assert _simulate(world, 12, seed) == 0


def test_model_detects_a_missing_wait():
    """sanity of the model itself: with ONE buffer (parity always 0) a fast rank does overwrite a slow reader"""
    import random

    def one_buffer(world, steps, seed):
        rnd = random.Random(seed)
        flags = [[0] * world for _ in range(world)]
        reading = [False] * world
        bad = 0
        progs = []
        for r in range(world):
            ops = []
            for c in range(1, steps + 1):
                peers = [p for p in range(world) if p != r]
                ops += [("data", c, p) for p in peers] + [("flag", c, p) for p in peers] + [("wait", c, None), ("begin", c, None), ("end", c, None)]
            progs.append(ops)
        pc = [0] * world
        while any(pc[r] < len(progs[r]) for r in range(world)):
            runnable = [r for r in range(world) if pc[r] < len(progs[r]) and not (
                progs[r][pc[r]][0] == "wait" and any(flags[r][q] < progs[r][pc[r]][1] for q in range(world) if q != r))]
            r = rnd.choice(runnable)
            kind, c, p = progs[r][pc[r]]
            pc[r] += 1
            if kind == "data" and reading[p]:
                bad += 1
            elif kind == "flag":
                flags[p][r] = c
            elif kind == "begin":
                reading[r] = True
            elif kind == "end":
                reading[r] = False
        return bad

    assert sum(one_buffer(3, 12, s) for s in range(40)) > 0


# ---- the mailbox query: why a CHANGE of root needs the previous root's call to have returned (yume_pdq_cuda.h) -------
def _simulate_mailbox(world: int, roots, seed: int, barrier_on_root_change: bool) -> int:
    """Hosts enqueue one kernel per query; a non-root's call returns at once, the root's when every rank has answered.
    GPU streams run their kernels in order:  [root: needles -> every mailbox, then the flag]  wait for the flag  scan
    (reads its own mailbox from begin to end)  answer -> root.  Returns how often needles were written into a mailbox
    that its owner was still scanning from (or had not even read yet)."""
    import random
    rnd = random.Random(seed)
    n_q = len(roots)
    flag = [0] * world                      # seq_in of every mailbox
    answers = [set() for _ in range(n_q + 1)]
    scanning = [False] * world
    host_pc = [1] * world                   # next query the host of rank r will enqueue
    returned = [0] * world                  # last query whose call has returned on rank r
    gpu = [[] for _ in range(world)]        # enqueued, not yet executed GPU events
    bad = 0

    def enqueue(r, seq):
        root = roots[seq - 1]
        ops = []
        if r == root:
            ops += [("data", seq, p) for p in range(world)] + [("flag", seq, p) for p in range(world)]
        ops += [("wait", seq, None), ("begin", seq, None), ("end", seq, None), ("answer", seq, root)]
        gpu[r] += ops

    while any(returned[r] < n_q for r in range(world)):
        moves = []
        for r in range(world):
            seq = host_pc[r]
            if seq <= n_q and returned[r] == seq - 1:
                changed = seq > 1 and roots[seq - 1] != roots[seq - 2]
                if not (barrier_on_root_change and changed) or all(returned[x] >= seq - 1 for x in range(world)):
                    moves.append(("enqueue", r))
            if seq <= n_q + 1 and returned[r] < seq - 1:
                # the call of query seq - 1 is still open: only the root's blocks
                q = seq - 1
                if roots[q - 1] != r or len(answers[q]) == world:
                    moves.append(("return", r))
            if gpu[r]:
                kind, seq_g, p = gpu[r][0]
                if not (kind == "wait" and flag[r] < seq_g):
                    moves.append(("gpu", r))
        if not moves:
            return bad + 1000               # a flag went backwards (two roots' stores crossed): the real kernels time out
        what, r = rnd.choice(moves)
        if what == "enqueue":
            enqueue(r, host_pc[r])
            host_pc[r] += 1
        elif what == "return":
            returned[r] = host_pc[r] - 1
        else:
            kind, seq_g, p = gpu[r].pop(0)
            if kind == "data" and scanning[p]:
                bad += 1
            elif kind == "wait" and flag[r] != seq_g:
                bad += 1                    # the next query's needles are already here (the real kernel would time out)
            elif kind == "flag":
                flag[p] = seq_g
            elif kind == "begin":
                scanning[r] = True
            elif kind == "end":
                scanning[r] = False
            elif kind == "answer":
                answers[seq_g].add(r)
    return bad


def test_mailbox_root_change_needs_the_previous_call_to_have_returned():
    same_root = [0] * 8
    changing = [0, 1, 2, 0, 2, 1, 1, 0]
    assert sum(_simulate_mailbox(3, same_root, s, False) for s in range(60)) == 0      # one root: nothing to add
    # a new root posts as soon as ITS OWN previous kernel is done: the old root (or any third rank) may still be scanning
    assert sum(_simulate_mailbox(2, [0, 1] * 4, s, False) for s in range(60)) > 0
    assert sum(_simulate_mailbox(3, changing, s, False) for s in range(200)) > 0
    # ShardedMatcher.scan_from's barrier when the root changes: the previous root's call has returned = everyone is done
    assert sum(_simulate_mailbox(2, [0, 1] * 4, s, True) for s in range(60)) == 0
    assert sum(_simulate_mailbox(3, changing, s, True) for s in range(200)) == 0
    assert sum(_simulate_mailbox(8, [i % 8 for i in range(10)], s, True) for s in range(20)) == 0
