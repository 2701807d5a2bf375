"""Multi-GPU plumbing for the hash-and-match path: one process per GPU over torch.distributed.

The reference has no distributed layer at all; this is the only new
"exchange" surface and it is deliberately tiny because both halves of the path shard into
independent units:

  * frames are split by batch (rank r hashes its own frames),
  * the database is sharded by contiguous row range (rank r holds rows
    [row_begin(r), row_begin(r+1)) resident in its HBM and reports GLOBAL row indices),
  * exchange 1: all-gather of the needle hashes (256 B per image),
  * exchange 2: gather of the per-rank match lists (16 B per match) + merge in
    (row, needle) order -- the order CpuMatcher::find reports (matching/mod.rs:336-362).

No reduction, no all-to-all.  Works with the NCCL backend on GPUs and with gloo on CPU
(the host logic is covered by world_size-2 gloo tests; the scan itself has no CPU path).
"""
from __future__ import annotations

import numpy as np

MATCH_DTYPE = np.dtype([("row", "<u8"), ("needle", "<u4"), ("dist", "<u4")])


def shard_range(n_rows: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous row range [begin, end) of `rank`; the last rank takes the remainder."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    per = n_rows // world
    begin = rank * per
    end = n_rows if rank == world - 1 else begin + per
    return begin, end


def needle_base(rank: int, images_per_rank: int, dihedrals: int = 8) -> int:
    """Index of rank's first needle in the all-gathered needle array."""
    return rank * images_per_rank * dihedrals


def merge_match_lists(lists) -> np.ndarray:
    """Concatenate per-rank match arrays (MATCH_DTYPE) and sort by (row, needle)."""
    lists = [np.asarray(x, dtype=MATCH_DTYPE).reshape(-1) for x in lists]
    allm = np.concatenate(lists) if lists else np.zeros(0, dtype=MATCH_DTYPE)
    order = np.lexsort((allm["needle"], allm["row"]))
    return allm[order]


def pack_matches(m: np.ndarray, cap: int):
    """Fixed-capacity wire format for the gather: (count int64[1], payload uint8[cap, 16])."""
    m = np.asarray(m, dtype=MATCH_DTYPE).reshape(-1)
    payload = np.zeros((cap, 16), dtype=np.uint8)
    k = min(len(m), cap)
    payload[:k] = m[:k].view(np.uint8).reshape(-1, 16)
    return np.array([len(m)], dtype=np.int64), payload


def unpack_matches(count: int, payload: np.ndarray) -> np.ndarray:
    k = min(int(count), payload.shape[0])
    return np.ascontiguousarray(payload[:k]).view(MATCH_DTYPE).reshape(-1).copy()


def all_gather_needles(local_needles, group=None):
    """All-gather [q_local, 32] uint8 needle tensors (same q_local on every rank) -> [world*q_local, 32]."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world * local_needles.shape[0], 32), dtype=torch.uint8, device=local_needles.device)
    dist.all_gather_into_tensor(out, local_needles.contiguous(), group=group)
    return out


def gather_match_lists(local_matches: np.ndarray, cap: int, device=None, group=None) -> np.ndarray:
    """Gather every rank's match list on all ranks and merge in (row, needle) order.
    Raises if any rank had more than `cap` matches (caller re-runs with a larger cap)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    cnt, payload = pack_matches(local_matches, cap)
    dev = device if device is not None else "cpu"
    t_cnt = torch.from_numpy(cnt).to(dev)
    t_pay = torch.from_numpy(payload).to(dev)
    all_cnt = torch.empty(world, dtype=torch.int64, device=dev)
    all_pay = torch.empty((world * cap, 16), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(all_cnt, t_cnt, group=group)
    dist.all_gather_into_tensor(all_pay, t_pay, group=group)
    counts = all_cnt.cpu().numpy()
    if (counts > cap).any():
        raise OverflowError(f"a rank produced {int(counts.max())} matches, capacity {cap}")
    pay = all_pay.cpu().numpy().reshape(world, cap, 16)
    return merge_match_lists([unpack_matches(counts[r], pay[r]) for r in range(world)])


class ShardedMatcher:
    """Row-range sharded `CudaMatcher` over a torch.distributed process group (one rank per GPU).

    ``rows_factory(begin, end, device)`` must return a ``CudaVectorDatabase`` holding global rows
    [begin, end) with ``global_row_offset=begin`` (e.g. uploaded from a file slice, or synthetic)."""

    def __init__(self, n_rows: int, threshold: int, rows_factory, device: int, group=None):
        import torch.distributed as dist
        from . import CudaMatcher
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.device = device
        self.begin, self.end = shard_range(n_rows, self.rank, self.world)
        self.db = rows_factory(self.begin, self.end, device)
        self.matcher = CudaMatcher(self.db, threshold)

    def find_all(self, local_needles, cap: int = 1 << 16) -> np.ndarray:
        """local_needles: this rank's [q_local, 32] uint8 CUDA tensor.  Returns the merged global match
        list (needle indices refer to the all-gathered needle array) on every rank."""
        if self.world == 1:
            return self.matcher.find_all(local_needles, cap=cap)
        needles = all_gather_needles(local_needles, self.group)
        local = self.matcher.find_all(needles, cap=cap)
        import torch
        return gather_match_lists(local, cap, device=torch.device("cuda", self.device), group=self.group)

    def scan(self, local_needles) -> bool:
        if self.world == 1:
            return self.matcher.scan(local_needles)
        import torch
        import torch.distributed as dist
        needles = all_gather_needles(local_needles, self.group)
        flag = torch.tensor([int(self.matcher.scan(needles))], dtype=torch.int32,
                            device=torch.device("cuda", self.device))
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)   # host-side OR of the flags
        return bool(flag.item())

    def close(self):
        self.db.close()
