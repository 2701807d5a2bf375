"""Two-GPU parity of the sharded pipeline (NCCL, one process per GPU): ShardedMatcher.screen /
find_all / scan against the CPU oracle on the unsharded problem.  Skipped on a single-GPU box (the
host-side logic is covered on CPU by tests/test_sharded_gloo.py)."""
import os
import socket

import numpy as np
import pytest

from helpers import flip_bits, match_tuples

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_local, n_rows, result_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import yume_pdq_b200 as yp
    from oracle import pyoracle
    from yume_pdq_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    orc = pyoracle.get()
    # every rank derives the same global problem with the oracle
    frames_all = orc.synth_frames(world * n_local, first_frame=500, kind=1)
    wh, _, _, _ = orc.hash_all_dihedrals(frames_all)
    needles_all = wh.reshape(-1, 32)
    rng = np.random.default_rng(17)
    rows = orc.synth_db(n_rows, first_row=3)
    for r in rng.choice(n_rows, size=120, replace=False):
        rows[r] = flip_bits(needles_all[int(rng.integers(0, needles_all.shape[0]))], int(rng.integers(0, 35)), rng)
    want = match_tuples(orc.matcher_find(rows, needles_all, 31))
    assert len(want) > 60

    def factory(begin, end, device):
        return yp.CudaVectorDatabase(rows[begin:end], device=device, global_row_offset=begin)

    sm = sharded.ShardedMatcher(n_rows, 31, factory, device=rank)
    mine = frames_all[rank * n_local:(rank + 1) * n_local]
    # host frames (pinned) through the chunked pipeline, partial last chunk
    pinned = torch.from_numpy(mine.copy()).pin_memory()
    hd = torch.zeros((n_local, 8, 32), dtype=torch.uint8, device=f"cuda:{rank}")
    got = sm.screen(pinned, hashes_out=hd, chunk=64)
    assert np.array_equal(hd.cpu().numpy(), wh[rank * n_local:(rank + 1) * n_local])
    assert match_tuples(got) == want
    # device frames, capacity growth
    got2 = sm.screen(torch.from_numpy(mine).cuda(), cap=8, chunk=64)
    assert match_tuples(got2) == want
    # the two-call path: hash, all-gather, find
    assert match_tuples(sm.find_all(hd.reshape(-1, 32))) == want
    assert sm.scan(hd.reshape(-1, 32)) is True
    # the small-needle query over peer memory (CUDA IPC mailboxes, no collective): which shards match?
    per = n_rows // world
    rr = np.random.default_rng(5)
    for trial in range(6):
        root = trial % world
        q = [8, 1, 32, 13, 8, 16][trial]
        needles = rr.integers(0, 256, size=(q, 32), dtype=np.uint8)
        if trial == 1:
            needles[0] = flip_bits(rows[per + 7], 31, rr)            # a row of rank 1's shard, exactly at the threshold
        if trial == 2:
            needles[5] = flip_bits(rows[3], 10, rr)                  # both shards
            needles[31] = flip_bits(rows[n_rows - 1], 0, rr)
        if trial == 3:
            needles[12] = flip_bits(rows[per - 1], 32, rr)           # one bit too far: no match
        if trial == 4:
            needles[:] = needles_all[:8]                             # this job's own hashes: planted rows exist somewhere
        want_mask = 0
        om = orc.matcher_find(rows, needles, 31)
        for r in om["row"]:
            want_mask |= 1 << min(int(r) // per, world - 1)
        src = needles if trial % 2 == 0 else torch.from_numpy(needles).cuda()     # host array or device tensor
        got_mask = sm.scan_from(src if rank == root else None, q, root=root)
        assert got_mask == (want_mask if rank == root else 0), (trial, got_mask, want_mask)
    assert trial == 5
    # many queries back to back, alternating roots: sequence numbers, mailbox reuse, no rank runs ahead of a scan
    fixed = [rr.integers(0, 256, size=(8, 32), dtype=np.uint8) for _ in range(4)]
    fixed[1][2] = flip_bits(rows[per + 11], 5, rr)
    fixed[2][7] = flip_bits(rows[17], 31, rr)
    fixed[3][0] = flip_bits(rows[per - 2], 3, rr)
    fixed[3][1] = flip_bits(rows[n_rows - 3], 30, rr)
    masks = []
    for nd in fixed:
        mk = 0
        for r in orc.matcher_find(rows, nd, 31)["row"]:
            mk |= 1 << min(int(r) // per, world - 1)
        masks.append(mk)
    assert masks == [0, 2, 1, 3]
    for it in range(200):
        k, root = it % 4, (it // 3) % world
        got_mask = sm.scan_from(fixed[k] if rank == root else None, 8, root=root)
        assert got_mask == (masks[k] if rank == root else 0), (it, got_mask)
    sm.close()
    np.save(os.path.join(result_dir, f"ok{rank}.npy"), np.array([len(got)]))
    dist.destroy_process_group()


def test_two_gpu_sharded_screen(tmp_path, oracle):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), 150, 60_001, str(tmp_path)), nprocs=world, join=True)
    counts = [int(np.load(tmp_path / f"ok{r}.npy")[0]) for r in range(world)]
    assert counts[0] == counts[1] > 60


def _fused_worker(rank, world, port, n_rows, result_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import yume_pdq_b200 as yp
    from oracle import pyoracle
    from yume_pdq_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    orc = pyoracle.get()
    rows = orc.synth_db(n_rows, first_row=3)
    per = n_rows // world

    def factory(begin, end, device):
        return yp.CudaVectorDatabase(rows[begin:end], device=device, global_row_offset=begin)

    sm = sharded.ShardedMatcher(n_rows, 31, factory, device=rank)
    rr = np.random.default_rng(23)
    queries = []
    for k in range(8):
        q = [8, 1, 32, 13, 8, 16, 8, 8][k]
        needles = rr.integers(0, 256, size=(q, 32), dtype=np.uint8)
        if k == 1:
            needles[0] = flip_bits(rows[per + 7], 31, rr)            # rank 1's shard, exactly at the threshold
        if k == 2:
            needles[5] = flip_bits(rows[3], 10, rr)                  # both shards
            needles[31] = flip_bits(rows[n_rows - 1], 0, rr)
        if k == 3:
            needles[12] = flip_bits(rows[per - 1], 32, rr)           # one bit too far: no match
        if k == 4:
            needles[7] = flip_bits(rows[per - 1], 30, rr)            # last row of rank 0's shard
        if k == 6:
            needles[0] = flip_bits(rows[per], 1, rr)                 # first row of rank 1's shard
        mk = 0
        for r in orc.matcher_find(rows, needles, 31)["row"]:
            mk |= 1 << min(int(r) // per, world - 1)
        queries.append((q, needles, mk))
    assert [m for _, _, m in queries] == [0, 2, 3, 0, 1, 0, 2, 0]
    assert sm.scan_from(queries[0][1] if rank == 0 else None, 8, root=0) == 0        # sets the mailboxes up
    # shards of >= 1M rows: AUTO is the fused form (one launch per rank); then both forced forms, then MIXED forms --
    # rank 0 fused, rank 1 separate launches and the other way round: one flag protocol
    for mode_of in (lambda r: yp.MAILBOX_AUTO, lambda r: yp.MAILBOX_FUSED, lambda r: yp.MAILBOX_SEPARATE,
                    lambda r: yp.MAILBOX_FUSED if r == 0 else yp.MAILBOX_SEPARATE,
                    lambda r: yp.MAILBOX_SEPARATE if r == 0 else yp.MAILBOX_FUSED):
        sm._mailbox.set_mode(mode_of(rank))
        for k, (q, needles, mk) in enumerate(queries):
            root = k % world
            src = needles if k % 3 else torch.from_numpy(needles).cuda()
            got = sm.scan_from(src if rank == root else None, q, root=root)
            assert got == (mk if rank == root else 0), (k, got, mk)
    # back to back, alternating roots and forms
    for it in range(300):
        k, root = it % 8, (it // 3) % world
        sm._mailbox.set_mode(yp.MAILBOX_FUSED if (it // 5 + rank) % 3 else yp.MAILBOX_SEPARATE)
        q, needles, mk = queries[k]
        got = sm.scan_from(needles if rank == root else None, q, root=root)
        assert got == (mk if rank == root else 0), (it, got, mk)
    torch.cuda.synchronize()
    sm.close()
    np.save(os.path.join(result_dir, f"fused_ok{rank}.npy"), np.array([1]))
    dist.destroy_process_group()


def test_two_gpu_fused_mailbox_query(tmp_path, oracle):
    """The one-launch-per-rank form of the peer-memory query (rows-in-M scan + MailboxHook) on two 1.1M-row shards."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    world = 2
    mp.spawn(_fused_worker, args=(world, _free_port(), 2_200_006, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"fused_ok{r}.npy") for r in range(world))


def _exchange_worker(rank, world, port, result_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import yume_pdq_b200 as yp
    from oracle import pyoracle
    from yume_pdq_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    orc = pyoracle.get()
    q_loc = 1024
    ex = sharded.make_exchange(rank, rank, world, q_loc * 32)
    n_rows = 300_000
    rows = orc.synth_db(n_rows, first_row=9)
    per = n_rows // world
    db = yp.CudaVectorDatabase(rows[rank * per:(rank + 1) * per if rank < world - 1 else n_rows], device=rank,
                               global_row_offset=rank * per)
    m = yp.CudaMatcher(db, 31)
    side = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    out = torch.zeros((1 << 12, 16), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    mine = torch.empty((q_loc, 32), dtype=torch.uint8, device="cuda")
    ev = torch.cuda.Event()
    for step in range(12):
        # every rank derives every rank's needles of this step; plants make some of them match rows of either shard
        rr = np.random.default_rng(1000 + step)
        needles_all = rr.integers(0, 256, size=(world * q_loc, 32), dtype=np.uint8)
        for k in range(40):
            needles_all[int(rr.integers(0, world * q_loc))] = flip_bits(rows[int(rr.integers(0, n_rows))], int(rr.integers(0, 32)), rr)
        main.wait_event(ev)                      # the previous push of `mine` is done
        mine.copy_(torch.from_numpy(needles_all[rank * q_loc:(rank + 1) * q_loc]))
        done = torch.cuda.Event()
        done.record(main)
        with torch.cuda.stream(side):
            side.wait_event(done)
            ex.post(mine, stream=side, how=yp.EXCHANGE_KERNEL if step % 3 else yp.EXCHANGE_COPY_ENGINE)
            ev.record(side)
        cnt.zero_()
        m.find_device(mine, q_loc, out, 1 << 12, cnt, needle_base=rank * q_loc)
        others = ex.wait(stream=main)
        m.find_device(others, (world - 1) * q_loc, out, 1 << 12, cnt, split_at=rank * q_loc, gap=q_loc)
        torch.cuda.synchronize()
        got = sorted(match_tuples(out[: int(cnt.item())].cpu().numpy().view(yp.MATCH_DTYPE).reshape(-1)))
        lo, hi = rank * per, ((rank + 1) * per if rank < world - 1 else n_rows)
        want = sorted(t for t in match_tuples(orc.matcher_find(rows, needles_all, 31)) if lo <= t[0] < hi)
        assert got == want, (step, len(got), len(want))
        assert len(want) > 5
        dist.barrier()                           # ranks stay within one exchange of each other
    assert not ex.failed()
    ex.close()
    db.close()
    np.save(os.path.join(result_dir, f"ex_ok{rank}.npy"), np.array([1]))
    dist.destroy_process_group()


def test_two_gpu_peer_exchange(tmp_path, oracle):
    """The copy-engine needle all-gather (yume_pdq_cuda_exchange_*) + find_device_split: twelve steps with fresh needles,
    every rank's match list must be the oracle's for its shard with the un-packed needle indices."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    world = 2
    mp.spawn(_exchange_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ex_ok{r}.npy") for r in range(world))
