"""Host-side multi-GPU logic on CPU: world_size-2 gloo process group.  The scan itself has no CPU
path, so each rank produces its local match list with the ORACLE on its row-range shard (the oracle
is the checker here, and the thing under test is the sharding / gather / merge plumbing)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from helpers import match_tuples, plant


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rows, q_local, result_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    from oracle import pyoracle
    from yume_pdq_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = pyoracle.get()
    # every rank derives the same global problem
    rng = np.random.default_rng(99)
    rows = orc.synth_db(n_rows)
    needles_all = rng.integers(0, 256, size=(world * q_local, 32), dtype=np.uint8)
    plant(rows, needles_all, [0, 31, 32], rng)
    begin, end = sharded.shard_range(n_rows, rank, world)
    # exchange 1: all-gather of the needles each rank "hashed"
    mine = torch.from_numpy(needles_all[rank * q_local:(rank + 1) * q_local].copy())
    gathered = sharded.all_gather_needles(mine).numpy()
    assert np.array_equal(gathered, needles_all)
    assert sharded.needle_base(rank, q_local // 8) == rank * q_local
    # local scan of the shard (oracle stands in for the device kernel), global row indices
    local = orc.matcher_find(rows[begin:end], gathered, 31)
    local["row"] += begin
    # exchange 2: gather + merge of match lists
    merged = sharded.gather_match_lists(local, cap=4096)
    want = orc.matcher_find(rows, needles_all, 31)
    assert match_tuples(merged) == match_tuples(want)
    # overflow is reported, not truncated silently
    try:
        sharded.gather_match_lists(local, cap=1)
        raise AssertionError("expected OverflowError")
    except OverflowError:
        pass
    np.save(os.path.join(result_dir, f"ok{rank}.npy"), np.array([len(merged)]))
    dist.destroy_process_group()


def test_two_rank_gather_merge(tmp_path, oracle):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, 20_001, 16, str(tmp_path)), nprocs=world, join=True)
    counts = [int(np.load(tmp_path / f"ok{r}.npy")[0]) for r in range(world)]
    assert counts[0] == counts[1] == 2 * 16 * 2


def test_shard_range_partition():
    from yume_pdq_b200 import sharded
    for n, w in [(10_000_384, 8), (10_000_384, 1), (1_000_000_000, 8), (7, 3), (0, 2)]:
        spans = [sharded.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    with pytest.raises(ValueError):
        sharded.shard_range(10, 2, 2)


def test_pack_merge_roundtrip():
    from yume_pdq_b200 import sharded
    rng = np.random.default_rng(0)
    a = np.zeros(5, dtype=sharded.MATCH_DTYPE)
    a["row"] = [9, 3, 3, 7, 1]
    a["needle"] = [0, 5, 2, 1, 8]
    a["dist"] = rng.integers(0, 32, 5)
    cnt, pay = sharded.pack_matches(a, 8)
    back = sharded.unpack_matches(cnt[0], pay)
    assert (back == a).all()
    m = sharded.merge_match_lists([a[:2], a[2:], a[:0]])
    assert [(int(r), int(n)) for r, n in zip(m["row"], m["needle"])] == [(1, 8), (3, 2), (3, 5), (7, 1), (9, 0)]
