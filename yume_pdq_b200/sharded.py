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


def make_exchange(device: int, rank: int, world: int, slice_bytes: int, group=None):
    """A connected PeerExchange (the copy-engine all-gather of every rank's needle slice): allocates this rank's
    buffers, all-gathers the 64-byte CUDA IPC handles over torch.distributed and maps the peers'.  COLLECTIVE."""
    import torch.distributed as dist
    from . import PeerExchange
    ex = PeerExchange(device, rank, world, slice_bytes)
    handles = [None] * world
    dist.all_gather_object(handles, ex.handle, group=group)
    ex.connect(handles)
    dist.barrier(group=group)
    return ex


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
        self._mailbox = None
        self._last_root = None      # root of the previous scan_from query
        self._exchange = None       # (chunk, PeerExchange): the needle exchange of screen()

    def find_all(self, local_needles, cap: int = 1 << 16) -> np.ndarray:
        """local_needles: this rank's [q_local, 32] uint8 CUDA tensor.  Returns the merged global match
        list (needle indices refer to the all-gathered needle array) on every rank."""
        if self.world == 1:
            return self.matcher.find_all(local_needles, cap=cap)
        needles = all_gather_needles(local_needles, self.group)
        local = self.matcher.find_all(needles, cap=cap)
        import torch
        return gather_match_lists(local, cap, device=torch.device("cuda", self.device), group=self.group)

    def screen(self, local_frames, hashes_out=None, cap: int = 1 << 16, chunk: int = 128,
               pixel_format: int = 0) -> np.ndarray:
        """Pipelined multi-GPU hash + match (BASELINE config 4: video screening against a sharded DB).

        ``local_frames``: this rank's [n_local, 512, 512] f32 frames (or 8-bit frames in ``pixel_format``:
        u8 [n, 512, 512], [n, 512, 512, 3] or [n, 512, 512, 4]) -- a (pinned) host numpy array / CPU tensor,
        or a CUDA tensor; every rank passes the same n_local.  Per chunk of ``chunk`` frames:
        [H2D of chunk i+1 on a copy stream] || [hash chunk i -> push its 8 * chunk needles into every peer's HBM
        (PeerExchange: one kernel over NVLink, no collective library) -> match this rank's own needles -> wait for the
        peers' -> match them in one launch (find_device_split), all against this rank's row shard].  One host
        synchronisation at the end, then the gather + merge of the match lists.  Returns, on every rank,
        all matches sorted by (row, needle) with needle = 8 * global_frame + dihedral and
        global_frame = rank * n_local + local index.  ``hashes_out`` ([n_local, 8, 32] uint8 CUDA tensor)
        receives this rank's hashes if given."""
        import torch
        import torch.distributed as dist
        from . import CudaKernel, MATCH_DTYPE as MD
        if self.world == 1:
            return self.matcher.screen(local_frames, hashes=hashes_out, cap=cap, pixel_format=pixel_format)
        dev = torch.device("cuda", self.device)
        n = int(local_frames.shape[0])
        W = self.world
        on_host = not (hasattr(local_frames, "is_cuda") and local_frames.is_cuda)
        if on_host and not hasattr(local_frames, "is_cuda"):
            local_frames = torch.from_numpy(np.ascontiguousarray(local_frames))
        main = torch.cuda.current_stream(dev)
        # every kernel of this call runs on the caller's current stream (PyTorch side streams do not synchronise
        # with the legacy default stream, so launching on stream 0 under torch.cuda.stream(...) would race)
        kern = CudaKernel(self.device, stream=main)
        hashes = hashes_out if hashes_out is not None else torch.empty((n, 8, 32), dtype=torch.uint8, device=dev)
        if self._exchange is None or self._exchange[0] != chunk:
            if self._exchange is not None:
                self._exchange[1].close()
            self._exchange = (chunk, make_exchange(self.device, self.rank, W, chunk * 8 * 32, group=self.group))
        ex = self._exchange[1]
        out = torch.zeros((cap, 16), dtype=torch.uint8, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        if on_host:
            copy = torch.cuda.Stream(dev)
            bufs = [torch.empty((chunk,) + tuple(local_frames.shape[1:]), dtype=local_frames.dtype, device=dev)
                    for _ in range(2)]
            copy.wait_stream(main)     # the staging buffers come from `main`'s allocator pool
            ev_copied = [torch.cuda.Event() for _ in range(2)]
            ev_done = [torch.cuda.Event() for _ in range(2)]
        n_chunks = (n + chunk - 1) // chunk
        for c in range(n_chunks):
            f0, k = c * chunk, min(chunk, n - c * chunk)
            b = c & 1
            if on_host:
                with torch.cuda.stream(copy):
                    if c >= 2:
                        copy.wait_event(ev_done[b])
                    bufs[b][:k].copy_(local_frames[f0:f0 + k], non_blocking=True)
                    ev_copied[b].record(copy)
                main.wait_event(ev_copied[b])
                src = bufs[b][:k]
            else:
                src = local_frames[f0:f0 + k]
            kern.hash_batch(src, want_quality=False, want_thresholds=False, out={"hashes": hashes[f0:f0 + k]},
                            pixel_format=pixel_format)
            if on_host:
                ev_done[b].record(main)
            # chunk-local needle index p = r * 8k + i (rank r, needle i of its chunk); rebased below to 8 * (r * n + f0) + i.
            # A rank can never be two exchanges ahead of a peer: its push of chunk c + 2 follows its wait for chunk c + 1,
            # which needs the peer's push of c + 1, which follows the peer's last read of chunk c (stream order).
            mine = hashes[f0:f0 + k].reshape(-1, 32)
            ex.post(mine, stream=main, nbytes=k * 8 * 32)
            base = c * W * chunk * 8
            self.matcher.find_device(mine, k * 8, out, cap, cnt, needle_base=base + self.rank * k * 8, stream=main)
            others = ex.wait(stream=main)
            self.matcher.find_device(others, (W - 1) * k * 8, out, cap, cnt, needle_base=base, split_at=self.rank * k * 8,
                                     gap=k * 8, stream=main)
        # every rank must take the same decision about a retry (the retry runs collectives): agree on the
        # largest per-rank match count
        worst = cnt.to(torch.int64)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=self.group)
        total, worst = int(cnt.item()), int(worst.item())
        if worst > cap:
            return self.screen(local_frames, hashes_out=hashes_out, cap=worst, chunk=chunk, pixel_format=pixel_format)
        local = out[:total].cpu().numpy().view(MD).reshape(-1).copy()
        cidx = local["needle"] // (W * chunk * 8)
        p = local["needle"] % (W * chunk * 8)
        k8 = np.minimum(chunk, n - cidx * chunk).astype(np.int64) * 8
        r, i = p // k8, p % k8
        local["needle"] = ((r * n + cidx * chunk) * 8 + i).astype(np.uint32)
        return gather_match_lists(local, cap, device=dev, group=self.group)

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

    def scan_from(self, needles, q: int, root: int = 0) -> int:
        """The sharded small-needle query over peer memory (no collective): up to 32 needles held by `root` (host array
        or device tensor; the other ranks pass None), every rank scans its shard.  COLLECTIVE: every rank calls it.
        Returns, on the root, the bit mask of the ranks whose shard holds a match (0 elsewhere) -- the multi-GPU form of
        VulkanMatcher::scan (vulkan.rs:349-406).  The mailboxes are set up on first use (one all-gather of 64-byte
        CUDA IPC handles)."""
        if self.world == 1:
            return int(self.matcher.scan(needles))
        if self._mailbox is None:
            import torch.distributed as dist
            from . import PeerMailbox
            mb = PeerMailbox(self.device, self.rank, self.world)
            handles = [None] * self.world
            dist.all_gather_object(handles, mb.handle, group=self.group)
            mb.connect(handles)
            dist.barrier(group=self.group)
            self._mailbox = mb
        if self._last_root is not None and root != self._last_root:
            # A root returns only once every rank has answered, i.e. finished reading its mailbox -- but only the ROOT
            # knows that.  A different rank may therefore post the next query's needles only after the previous root's
            # call has returned: a slower rank (the previous root included) could otherwise still be scanning, and
            # verifying candidates against, the mailbox the new root overwrites.  Same-root sequences need nothing.
            import torch.distributed as dist
            dist.barrier(group=self.group)
        self._last_root = root
        return self._mailbox.scan(self.db, needles if self.rank == root else None, q, self.matcher.threshold, root)

    def close(self):
        if self._mailbox is not None:
            self._mailbox.close()
            self._mailbox = None
        if self._exchange is not None:
            self._exchange[1].close()
            self._exchange = None
        self.db.close()
