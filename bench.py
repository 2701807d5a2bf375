#!/usr/bin/env python3
"""bench.py -- the contract benchmark for the PDQ hash-and-match hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A *step* is one pass of the hot path over one batch of synthetic input, per GPU:
  hash 4096 device-resident 512x512 f32 frames into 8 dihedral 256-bit hashes each
  (BASELINE config 2), [N > 1: all-gather the 32768 x N needles over NCCL],
  scan them against this rank's row-range shard of the 10,000,384-row x 256-bit
  synthetic database at threshold 31 (BASELINE config 3's database; the reference
  bench size, benches/matcher.rs:18-24), compacting every match, [N > 1: gather the
  match lists].  metric = hash+match images/s over the whole job ("weak": frames per
  GPU fixed, total DB fixed and sharded, so pairs per GPU stay constant).

The JSON line also carries the two other figures of BASELINE.json's metric -- the
8-needle 10M-row scan in microseconds / DB GB/s and the hash-only rate -- each with
its roofline, a CPU baseline (the AVX2 hasher + AVX-512/u64 matcher port of the
reference, oracle/cpu_baseline.cpp) timed on a bounded sample, the end-to-end number
through the host-pointer C ABI, clocks sampled during the timed region and the number
of kernels this library launched, and (every N):
  "parity"   the sharded matcher's merged match list == the CPU oracle's on a
             sub-sampled problem (200,000 rows x ALL of the job's needles): full
             equality, (row, needle, distance);
  "config5"  BASELINE config 5: 125M synthetic rows per GPU (1e9 over 8 GPUs),
             1024 images x 8 dihedrals per batch matched in ONE 8192-needle launch;
  "config4"  BASELINE config 4: 65,536 device-generated frames (total) hashed and
             matched against the sharded 10M-row DB through ShardedMatcher.screen;
  "scan".sharded_e2e_us (N > 1)  needles on rank 0 -> broadcast -> N shard scans ->
             merged (OR) result on the host.

Only this file's cpu_baseline / --impl reference / parity legs touch oracle/ (as the
measured CPU reference and as the checker); the GPU arm's timed regions never do.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "hash+match images/s vs 10M DB"
UNIT = "images/s"
DB_ROWS = (10_000_000 // 2048 + 1) * 2048      # 10,000,384 (benches/matcher.rs:18-24)
FRAMES_PER_GPU = 4096
THRESHOLD = 31
SEED = 0x79756D652D706471
N_PLANT = 64
# DRAM bytes (read + written) per database row / per frame that ONE launch of each kernel moved under `ncu --set full`
# at the default sizes (tools/ncu_full.sh -> profiles/r02q/traffic.csv); None = not captured for this build
TRAFFIC_SOURCE = "profiles/r02q/traffic.csv (ncu --set full of this build's kernels, one GPU, default sizes)"
TRAFFIC_TENSOR_PER_ROW = 17.139       # match_tensor4x2_kernel<X2Large>, 32768 needles: 164.0 MB read + 7.4 MB written / 10,000,384 rows
TRAFFIC_HASH_PER_FRAME = 1048602.9    # hash_fused_kernel: 1236.2 MB read + 5.4 MB written / 1184 frames (algorithmic 1,048,896)
TRAFFIC_SCAN_PER_ROW = 17.249         # match_rows_m_kernel<N = 8>: 160.0 MB read + 12.5 MB written / 10,000,384 rows (algorithmic 16)


def env_int(name, default):
    return int(os.environ.get(name, default))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "MEASURED_PEAKS.json", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# --------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (C++ port of the
    AVX2 hasher + AVX-512 VPOPCNTDQ / u64 matcher, oracle/cpu_baseline.cpp) on all host threads.
    Each step = a bounded sample of the workload: hash S frames, match their 8*S needles against
    the full 10,000,384-row database."""
    if rank != 0:
        return
    from oracle import pyoracle
    pyoracle.build()
    orc = pyoracle.get()
    cpu = pyoracle.CpuBaseline()
    threads = cpu.hw_threads
    S = args.cpu_sample
    rows = orc.synth_db(DB_ROWS, 0, SEED)
    frames = orc.synth_frames(S, 0, SEED, 1)

    def step():
        h, q, t, _ = cpu.hash_batch(frames, threads=threads, want_pdqf=False)
        return cpu.match(rows, h.reshape(-1, 32), THRESHOLD, True, threads)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = S * args.steps / dt
    kind = "port"
    sample = (f"{S} frames hashed (AVX2 port) + {8 * S} needles vs the full {DB_ROWS}-row DB per step, "
              f"{'AVX-512 VPOPCNTDQ' if cpu.has_avx512_popcnt else 'u64 popcnt'} matcher, {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u64",
        "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(n_gpus):
    return {
        "workload": (f"hash {FRAMES_PER_GPU} synthetic 512x512 f32 frames per GPU (8 dihedral hashes each) + "
                     f"match all needles vs a {DB_ROWS}-row x 256-bit synthetic DB at threshold {THRESHOLD}"),
        "frames_per_gpu": FRAMES_PER_GPU, "db_rows": DB_ROWS, "threshold": THRESHOLD,
        "needles_per_step": FRAMES_PER_GPU * 8 * n_gpus,
        "sharding": f"frames by batch, DB by row range over {n_gpus} GPU(s)",
        "l2": "inputs larger than L2 (4 GiB of frames + the DB stream through the 126 MB L2 every step)",
    }


def cpu_baseline_leg(args):
    """Rank 0, N=1 only: the CPU port timed on a bounded sample (about 10-30 s)."""
    from oracle import pyoracle
    pyoracle.build()
    orc = pyoracle.get()
    cpu = pyoracle.CpuBaseline()
    threads = cpu.hw_threads
    S = args.cpu_sample
    rows = orc.synth_db(DB_ROWS, 0, SEED)
    frames = orc.synth_frames(S, 0, SEED, 1)
    t0 = time.perf_counter()
    h, _, _, _ = cpu.hash_batch(frames, threads=threads, want_pdqf=False)
    t1 = time.perf_counter()
    reps = 0
    while True:
        cpu.match(rows, h.reshape(-1, 32), THRESHOLD, True, threads)
        reps += 1
        if time.perf_counter() - t1 > 8.0 or reps >= 4:
            break
    t2 = time.perf_counter()
    per_step = (t1 - t0) + (t2 - t1) / reps
    # single-thread figures comparable to the reference's published ones
    t3 = time.perf_counter()
    cpu.hash_batch(frames[:16], threads=1, want_pdqf=False)
    t4 = time.perf_counter()
    cpu.match(rows, h.reshape(-1, 32)[:8], THRESHOLD, False, 1)
    t5 = time.perf_counter()
    return {
        "value": S / per_step, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": (f"{S} frames hashed + {8 * S} needles matched vs the full {DB_ROWS}-row DB, "
                   f"AVX2 hasher port + {'AVX-512 VPOPCNTDQ' if cpu.has_avx512_popcnt else 'u64 popcnt'} "
                   f"matcher port, {threads} threads"),
        "hash_us_per_frame_1thread": (t4 - t3) / 16 * 1e6,
        "scan_8_needles_ms_1thread": (t5 - t4) * 1e3,
        "hash_images_per_s_all_threads": S / (t1 - t0),
    }


# --------------------------------------------------------------------------
_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def plant_rows(rng, needles_host, n_plant, row_lo, row_hi, write_row, own_lo=None, own_hi=None, stride=1):
    """Plants copies of random needles (0..33 bits away) at random rows of [row_lo, row_hi) that are multiples of
    `stride`; write_row(global_row, bytes32) is called for rows inside [own_lo, own_hi).  Deterministic in rng, so
    every rank computes the same list.  Returns [(global_row, needle_index, distance)]."""
    planted = []
    slots = rng.choice((row_hi - row_lo + stride - 1) // stride, size=n_plant, replace=False)
    for slot in slots:
        row = row_lo + int(slot) * stride
        j = int(rng.integers(0, needles_host.shape[0]))
        d = int(rng.integers(0, 34))
        bits = np.unpackbits(needles_host[j].copy())
        bits[rng.choice(256, size=d, replace=False)] ^= 1
        if own_lo is None or own_lo <= row < own_hi:
            write_row(row, np.packbits(bits))
        planted.append((row, j, d))
    return planted


def main():
    global DB_ROWS, FRAMES_PER_GPU, _REAL_STDOUT
    # Libraries write banners to stdout (NCCL prints "NCCL version ..." at the first collective): keep fd 1 for the
    # JSON line alone and send everything else to stderr.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=256, help="frames per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison on the sub-sampled problem")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--no-worst-case", action="store_true", help="skip the clustered-haystack leg")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--nccl-gather", action="store_true",
                    help="N > 1: all-gather the needles with NCCL instead of the copy-engine exchange over peer memory")
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help=argparse.SUPPRESS)
    ap.add_argument("--db-rows", type=int, default=DB_ROWS, help=argparse.SUPPRESS)
    ap.add_argument("--c5-rows-per-gpu", type=int, default=125_000_000, help=argparse.SUPPRESS)
    ap.add_argument("--c4-frames", type=int, default=65_536, help=argparse.SUPPRESS)
    args = ap.parse_args()
    DB_ROWS, FRAMES_PER_GPU = args.db_rows, args.frames
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import yume_pdq_b200 as yp

    if not torch.cuda.is_available() or yp.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    B = args.frames
    hbm_peak, peak_src, peaks = measured_peaks()
    main_stream = torch.cuda.current_stream(dev)

    def barrier():
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if n_gpus > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- setup (untimed): inputs resident in HBM ----------------------------
    kern = yp.CudaKernel(local_rank)
    frames = torch.empty((B, 512, 512), dtype=torch.float32, device=dev)
    yp.synth_frames(frames, B, first_frame=rank * B, seed=SEED, kind=1, device=local_rank)
    from yume_pdq_b200 import sharded
    row0, row1 = sharded.shard_range(DB_ROWS, rank, n_gpus)
    shard_rows = row1 - row0
    sm = sharded.ShardedMatcher(
        DB_ROWS, THRESHOLD,
        lambda b, e, d: yp.CudaVectorDatabase(device=d, global_row_offset=b, synthetic_rows=e - b, seed=SEED),
        device=local_rank)
    db, matcher = sm.db, sm.matcher
    outs = {"hashes": torch.empty((B, 8, 32), dtype=torch.uint8, device=dev),
            "quality": torch.empty((B, 8), dtype=torch.float32, device=dev),
            "thresholds": torch.empty((B, 8), dtype=torch.float32, device=dev)}
    q_loc = B * 8
    q_all = q_loc * n_gpus
    local_needles = outs["hashes"].reshape(-1, 32)
    needles_all = torch.empty((q_all, 32), dtype=torch.uint8, device=dev) if n_gpus > 1 else local_needles
    # match list with its count in a header row: ONE all-gather moves both (N > 1)
    CAP = 1 << 16 if n_gpus == 1 else 1 << 12
    list_buf = torch.zeros((1 + CAP, 16), dtype=torch.uint8, device=dev)
    match_cnt = list_buf[0, :4].view(torch.int32)
    match_out = list_buf[1:]
    gathered = torch.empty((n_gpus, 1 + CAP, 16), dtype=torch.uint8, device=dev) if n_gpus > 1 else None
    side = torch.cuda.Stream(dev) if n_gpus > 1 else None
    ev_hash, ev_gath, ev_match, ev_list = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
    others = torch.empty((q_all - q_loc, 32), dtype=torch.uint8, device=dev) if n_gpus > 1 else None
    exchange = exchange_lists = None
    if n_gpus > 1 and not args.nccl_gather:
        exchange = sharded.make_exchange(local_rank, rank, n_gpus, q_loc * 32)
        exchange_lists = sharded.make_exchange(local_rank, rank, n_gpus, (1 + CAP) * 16)

    # plant rows so the scan has something to find: copies of this job's own hashes, 0..31 bits away
    kern.hash_batch(frames, out=outs)
    torch.cuda.synchronize()
    local_hashes = local_needles.cpu().numpy()
    rng = np.random.default_rng(1234 + rank)
    planted = []
    for k in range(N_PLANT):
        j = int(rng.integers(0, local_hashes.shape[0]))
        lrow = int(rng.integers(0, shard_rows))
        h = np.unpackbits(local_hashes[j].copy())
        flips = rng.choice(256, size=int(rng.integers(0, 32)), replace=False)
        h[flips] ^= 1
        db.write_rows(lrow, np.packbits(h)[None])
        planted.append((row0 + lrow, rank * q_loc + j))

    def step(tl=None):
        # tl: optional dict of timing events (the untimed timeline pass after the timed region)
        def mark(name, stream=None):
            if tl is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(stream or main_stream)
                tl[name] = e
        mark("start")
        kern.hash_batch(frames, out=outs)
        mark("hashed")
        if n_gpus == 1:
            match_cnt.zero_()
            matcher.find_device(local_needles, q_loc, match_out, CAP, match_cnt, algo=yp.ALGO_AUTO)
            return
        # N > 1.  Default: NO collective library inside the step -- the needles and the {count, match list} blocks travel
        # through the peer exchange (yume_pdq_cuda_exchange_*: one push kernel stores a rank's block into every peer's HBM
        # over NVLink, flag words order it, a one-warp kernel waits), everything on ONE stream:
        #   hash -> push needles -> match OWN needles -> wait for the peers' needles -> match them in ONE launch
        #   (find_device_split: the packed batch of the other ranks, indices of the un-packed list) -> push the list block;
        # the peers' list blocks of step i are waited for at the end of the timed region / by whoever consumes them.
        # --nccl-gather: the NCCL formulation (needle all-gather + two packing copies on a side stream beside the own-needle
        # match, list all-gather under the next hash); measured slower because NCCL's kernels share the SMs with the
        # persistent matcher and with the hash kernel (step_timeline)
        if exchange is not None:
            exchange.post(local_needles, stream=main_stream)
            mark("gathered")
            mark("packed")
            match_cnt.zero_()
            matcher.find_device(local_needles, q_loc, match_out, CAP, match_cnt, needle_base=rank * q_loc)
            mark("own_matched")
            others_ptr = exchange.wait(stream=main_stream)      # one-warp kernel: every peer's slice is in
            matcher.find_device(others_ptr, q_all - q_loc, match_out, CAP, match_cnt, split_at=rank * q_loc, gap=q_loc)
            mark("others_matched")
            exchange_lists.post(list_buf, stream=main_stream)
            mark("lists_gathered")
            return
        ev_hash.record(main_stream)
        with torch.cuda.stream(side):
            side.wait_event(ev_hash)
            dist.all_gather_into_tensor(needles_all, local_needles)
            mark("gathered", side)
            if rank > 0:
                others[: rank * q_loc].copy_(needles_all[: rank * q_loc], non_blocking=True)
            if rank < n_gpus - 1:
                others[rank * q_loc:].copy_(needles_all[(rank + 1) * q_loc:], non_blocking=True)
            mark("packed", side)
            ev_gath.record(side)
        main_stream.wait_event(ev_list)      # the previous step's list has been gathered before the counter is reset
        match_cnt.zero_()
        matcher.find_device(local_needles, q_loc, match_out, CAP, match_cnt, needle_base=rank * q_loc)
        mark("own_matched")
        main_stream.wait_event(ev_gath)
        matcher.find_device(others, q_all - q_loc, match_out, CAP, match_cnt, split_at=rank * q_loc, gap=q_loc)
        mark("others_matched")
        ev_match.record(main_stream)
        with torch.cuda.stream(side):
            side.wait_event(ev_match)
            dist.all_gather_into_tensor(gathered.reshape(-1, 16), list_buf)
            mark("lists_gathered", side)
            ev_list.record(side)

    def finish_lists():
        """The last step's match lists are on every rank (inside the timed region); exchange: copy them into `gathered`."""
        if n_gpus == 1:
            return
        if exchange is None:
            main_stream.wait_event(ev_list)
            return
        exchange_lists.wait(stream=main_stream)

    def collect_lists():
        """untimed: the exchange's buffer (the other ranks' blocks, rank order) + this rank's own block -> `gathered`"""
        if exchange is None:
            return
        ptr = exchange_lists.wait(stream=main_stream)
        blk = (1 + CAP) * 16
        peers = yp.device_view(ptr, (n_gpus - 1) * blk).reshape(n_gpus - 1, 1 + CAP, 16)
        for r in range(n_gpus):
            gathered[r].copy_(list_buf if r == rank else peers[r if r < rank else r - 1])
        torch.cuda.synchronize()

    def exchange_find_all(mt):
        """The timed step's sequence (push needles, own match, wait, split match, push list block) against matcher `mt`;
        returns the merged match list, sorted by (row, needle).  COLLECTIVE; used by the parity leg."""
        match_cnt.zero_()
        exchange.post(local_needles, stream=main_stream)
        mt.find_device(local_needles, q_loc, match_out, CAP, match_cnt, needle_base=rank * q_loc)
        ptr = exchange.wait(stream=main_stream)
        mt.find_device(ptr, q_all - q_loc, match_out, CAP, match_cnt, split_at=rank * q_loc, gap=q_loc)
        exchange_lists.post(list_buf, stream=main_stream)
        collect_lists()
        g = gathered.cpu().numpy()
        parts = []
        for r in range(n_gpus):
            n = int(g[r, 0, :4].view(np.int32)[0])
            if n > CAP:
                raise SystemExit(f"bench.py: parity: rank {r} reports {n} matches, list capacity {CAP}")
            parts.append(g[r, 1:1 + n].reshape(-1).view(yp.MATCH_DTYPE).reshape(-1))
        m = np.concatenate(parts)
        dist.barrier()
        return m[np.lexsort((m["needle"], m["row"]))]

    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region ----------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    launches0 = yp.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    finish_lists()      # the last step's match lists have arrived on every rank
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    launches = yp.launch_count() - launches0
    clocks = sampler.stop()
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if n_gpus > 1:
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    elapsed_ms = max_over_ranks(elapsed_ms)
    ms_per_step = elapsed_ms / args.steps
    value = n_gpus * B / (ms_per_step * 1e-3)

    # ---- untimed: where a step's time goes (CUDA events inside four more steps; ms after the step's start, median) ----
    timeline = None
    if n_gpus > 1:
        tls = [dict() for _ in range(5)]
        for tl in tls:
            step(tl)
        torch.cuda.synchronize()
        names = ["hashed", "gathered", "packed", "own_matched", "others_matched", "lists_gathered"]
        mine = {n: float(np.median([tl["start"].elapsed_time(tl[n]) for tl in tls[1:]])) for n in names}
        mine["next_start"] = float(np.median([tls[i]["start"].elapsed_time(tls[i + 1]["start"]) for i in range(1, 4)]))
        per_rank = [None] * n_gpus
        dist.all_gather_object(per_rank, {"own_match": mine["own_matched"] - mine["packed"],
                                          "others_match": mine["others_matched"] - mine["own_matched"],
                                          "step": mine["next_start"]})
        timeline = {"rank0_ms": mine, "max_over_ranks_ms": {n: max_over_ranks(v) for n, v in mine.items()},
                    "per_rank_ms": per_rank,
                    "note": ("one stream: hashed -> gathered = packed (needles pushed to every peer) -> own_matched -> [wait for "
                             "the peers' needles] others_matched -> lists_gathered (list block pushed to every peer)")
                    if exchange is not None else
                    ("main stream: hashed -> own_matched -> others_matched; side stream: gathered (NCCL all-gather of "
                     "the needles) -> packed (two D2D copies) and lists_gathered (NCCL all-gather of count + list, "
                     "under the next step's hash)")}
        barrier()

    # ---- correctness of what was timed (untimed check) -----------------------
    n_matches = int(match_cnt.item())
    if n_matches > CAP:
        raise SystemExit(f"bench.py: rank {rank}: {n_matches} matches overflow the list capacity {CAP}")
    got = match_out[:n_matches].cpu().numpy().view(yp.MATCH_DTYPE).reshape(-1)
    got_set = {(int(r), int(n)) for r, n in zip(got["row"], got["needle"])}
    missing = [p for p in planted if p not in got_set]
    if missing:
        raise SystemExit(f"bench.py: rank {rank}: {len(missing)} planted matches were not found")
    total_matches = n_matches
    if n_gpus > 1:
        collect_lists()
        g = gathered.cpu().numpy()
        total_matches = int(sum(int(g[r, 0, :4].view(np.int32)[0]) for r in range(n_gpus)))

    # ---- parity: merged multi-GPU match list == CPU oracle on a sub-sampled problem ------
    parity = None
    if not args.no_parity:
        parity = parity_leg(args, yp, sharded, dist, torch, dev, rank, n_gpus, local_rank, local_needles, needles_all, q_all,
                            exchange_find_all if exchange is not None else None)

    # ---- per-kernel timings for the rooflines (CUDA events on the launch stream) ------
    def time_kernel(fn, iters, flush=None):
        """CUDA events on the launch stream.  The whole sequence [flush, e0, kernel, e1] x iters is
        enqueued without host synchronisation, so the GPU never waits for the host between the
        event and the kernel (a host-paced loop adds ~10 us of launch latency to every sample)."""
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        if flush is not None:
            flush.sum()
        for a, b in ev:
            if flush is not None:
                flush.sum()   # READ 256 MB: evicts L2 without leaving dirty lines to write back under the timed kernel
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        ts = [a.elapsed_time(b) for a, b in ev]
        return float(np.mean(ts)), float(np.median(ts)), float(np.min(ts))

    # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the ncu --set full captures of THIS
    # round's kernels (profiles/r02q/traffic.csv, one GPU, default sizes), scaled per row / per frame; null elsewhere
    default_sizes = n_gpus == 1 and DB_ROWS == (10_000_000 // 2048 + 1) * 2048
    traffic_tensor = TRAFFIC_TENSOR_PER_ROW * shard_rows if default_sizes and B == 4096 and TRAFFIC_TENSOR_PER_ROW else None
    traffic_hash = TRAFFIC_HASH_PER_FRAME * B if default_sizes and TRAFFIC_HASH_PER_FRAME else None
    traffic_scan = TRAFFIC_SCAN_PER_ROW * shard_rows if default_sizes and TRAFFIC_SCAN_PER_ROW else None
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    line_extra = {}
    mxf4_peak, mxf4_clk = yp.measure_mxf4_peak(local_rank, 20000)
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)   # 256 MB > 126 MB L2
    nd = local_needles
    # (1) dominant kernel of the step: large-batch match on the tensor cores
    m_mean, m_med, _ = time_kernel(
        lambda: (match_cnt.zero_(), matcher.find_device(nd, q_loc, match_out, CAP, match_cnt)), 3)
    pairs = shard_rows * q_loc
    # (2) hash kernel (HBM bound): algorithmic bytes = frames in + hashes/quality/thresholds out
    h_mean, h_med, h_min = time_kernel(lambda: kern.hash_batch(frames, out=outs), 10, flush)
    hash_bytes = B * (512 * 512 * 4 + 8 * 32 + 8 * 4 + 8 * 4)
    # (3) headline scan: 8 needles vs the resident shard (HBM bound), L2 flushed between launches
    n8 = nd[:8].contiguous()
    s_mean, s_med, s_min = time_kernel(
        lambda: matcher.find_device(n8, 8, match_out, CAP, match_cnt), 20, flush)
    # the same scan back to back WITHOUT evicting L2: a shard whose prefix plane (16 B per row) fits the 126 MB L2 -- 10M
    # rows over 4 or 8 GPUs -- stays resident between queries in a serving loop; reported beside, never instead of, `us`
    l2_mean = None
    if shard_rows * 16 <= 100e6:
        l2_mean, _, _ = time_kernel(lambda: matcher.find_device(n8, 8, match_out, CAP, match_cnt), 20, None)
    rows_in_m = shard_rows >= 1_000_000          # ALGO_AUTO's rule for <= 8 needles at threshold <= 40 (matcher.cu)
    p_mean, _, _ = time_kernel(
        lambda: matcher.find_device(n8, 8, match_out, CAP, match_cnt, algo=yp.ALGO_TILED), 20, flush)
    # (3b) the needle counts of the reference's batched CpuMatcher<B, N> and of a 128-frame video chunk
    batch_us = {}
    for qb in (16, 32, 128, 1024):
        if qb <= q_loc:
            nb = nd[:qb].contiguous()
            b_mean, _, _ = time_kernel(lambda: matcher.find_device(nb, qb, match_out, CAP, match_cnt), 10, flush)
            batch_us[str(qb)] = b_mean * 1e3
    scan_bytes = shard_rows * 32
    share = m_mean / (m_mean + h_mean)
    # (3c) N > 1: the sharded 8-needle query end to end (needles on rank 0 -> broadcast -> N shard scans -> OR)
    sharded_e2e = sharded_p2p = sharded_p2p_root = p2p_stamps = p2p_forms = None
    if n_gpus > 1:
        q8 = n8.clone()
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        flush_small = flush

        def sharded_scan():
            dist.broadcast(q8, src=0)
            match_cnt.zero_()
            matcher.find_device(q8, 8, match_out, CAP, match_cnt)
            flag.copy_(match_cnt.clamp(max=1))
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            return bool(flag.item())     # host-visible answer: includes the synchronisation

        for _ in range(5):
            sharded_scan()
        barrier()
        ts = []
        for _ in range(30):
            t0 = time.perf_counter()
            sharded_scan()
            ts.append((time.perf_counter() - t0) * 1e6)
        sharded_e2e = max_over_ranks(float(np.median(ts)))

        # the same query over peer memory (yume_pdq_cuda_mailbox_*: CUDA IPC mailboxes, plain NVLink stores ordered by
        # flags, no collective): the needles start in HOST memory on rank 0, the answer ends as a host bit mask on rank 0
        host_q8 = n8.cpu().numpy()
        p2p_mask = None
        p2p_forms = {}
        # FUSED: post, wait, rows-in-M scan, publish and collect inside ONE launch per rank (matcher_tc1.cu);
        # SEPARATE: five launches on the root, three elsewhere, around ALGO_AUTO's kernel; AUTO (the default, reported
        # as sharded_p2p_us): fused where ALGO_AUTO's kernel for the shard is the rows-in-M scan (>= 1M rows)
        for form, mode in (("fused", yp.MAILBOX_FUSED), ("separate", yp.MAILBOX_SEPARATE), ("auto", yp.MAILBOX_AUTO)):
            for i in range(5):
                p2p_mask = sm.scan_from(host_q8 if rank == 0 else None, 8, root=0)
                if i == 0:
                    sm._mailbox.set_mode(mode)
            barrier()
            ts = []
            for _ in range(30):
                flush_small.sum()     # evict the shard's prefix plane from L2 as the flushed kernel timings do
                torch.cuda.synchronize()
                if n_gpus > 1:
                    dist.barrier()
                t0 = time.perf_counter()
                p2p_mask = sm.scan_from(host_q8 if rank == 0 else None, 8, root=0)
                ts.append((time.perf_counter() - t0) * 1e6)
            p2p_forms[form] = {"us": max_over_ranks(float(np.median(ts))), "us_on_root": float(np.median(ts))}
        sharded_p2p = p2p_forms["auto"]["us"]
        sharded_p2p_root = p2p_forms["auto"]["us_on_root"] if rank == 0 else None
        p2p_stamps = sm._mailbox.last_timing() if rank == 0 else None
    worst = None
    if rank == 0 and not args.no_worst_case:
        worst = worst_case_leg(args, yp, torch, dev, local_rank, shard_rows, flush)
    del flush

    if rank == 0:
        # Tensor roofline.  At threshold <= 36 the matcher runs the 128-bit prefix filter on the tensor cores with two
        # rows per accumulator column (match_tensor4x2_kernel): one (row, needle) pair = 128 multiply-adds = 256 ops of
        # the packed-FP4 MMA (tcgen05.mma kind::mxf4, e2m1 operands, f32 accumulate; in the high field 6 of the 128
        # element positions carry the constant that aligns the two s16 fields, 122 carry hash bits) plus an exact
        # POPC check of the rare candidates.  Above threshold 36 the full 256-bit dot product runs
        # (match_tensor4_kernel), 512 ops.  Peak: MEASURED_PEAKS.json holds a dense bf16 rate but no FP4 one, and 4 x
        # its bf16 figure (FP4 : bf16 = 4 : 1 on the tensor pipe) is BELOW what this kernel achieves (cuBLAS bf16 on
        # random data is power-limited to ~1.4 GHz, +-1 operands are not), so the denominator is SELF-MEASURED: the
        # issue rate of the very same instruction measured live on this GPU with no loads and no epilogue
        # (yume_pdq_cuda_measure_mxf4_peak); the bf16-derived figure is kept beside it.
        paired = THRESHOLD <= 36
        ops_per_pair = 256 if paired else 512
        bf16 = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
        achieved_tops = pairs * ops_per_pair / (m_mean * 1e-3) / 1e12
        line_extra["roofline"] = {
            "kernel": "%s (tcgen05.mma kind::mxf4; needle batch %d x this rank's %d-row shard)"
                      % ("match_tensor4x2_kernel: 128-bit prefix filter, two rows per f32 accumulator" if paired else
                         "match_tensor4_kernel<256-bit>", q_loc, shard_rows),
            "bound": "tensor", "achieved": achieved_tops, "peak": mxf4_peak, "unit": "TFLOP/s",
            "frac": achieved_tops / mxf4_peak,
            "peak_is_self_measured": True,
            "peak_source": ("SELF-MEASURED (MEASURED_PEAKS.json has no FP4 entry): back-to-back tcgen05.mma kind::mxf4 "
                            "M128 N240 K64 on this GPU, %.1f SM clk per MMA (yume_pdq_cuda_measure_mxf4_peak)" % mxf4_clk),
            "peak_bf16_x4": 4.0 * bf16, "frac_of_bf16_x4": achieved_tops / (4.0 * bf16),
            "peak_bf16_x4_source": "4 x bf16_tflops_sustained of MEASURED_PEAKS.json" if peaks else
                                   "4 x 1590 (B200_PROFILING.md fallback)",
            "tensor_ops_per_pair": ops_per_pair, "executed_tensor_ops_per_pair": ops_per_pair,
            "executed_frac": achieved_tops / mxf4_peak,
            "tpairs_per_s": pairs / (m_mean * 1e-3) / 1e12,
            "traffic": traffic_tensor, "traffic_source": TRAFFIC_SOURCE,
            "ms_per_launch": m_mean, "share_of_step": share,
        }
        line_extra["roofline_hash"] = {
            "kernel": "hash_fused_kernel", "bound": "hbm", "achieved": hash_bytes / (h_mean * 1e-3) / 1e9,
            "peak": hbm_peak, "unit": "GB/s", "frac": hash_bytes / (h_mean * 1e-3) / 1e9 / hbm_peak,
            "peak_source": peak_src, "traffic": traffic_hash, "algorithmic_bytes": hash_bytes,
            "traffic_source": TRAFFIC_SOURCE, "ms_per_launch": h_mean,
            "images_per_s": B / (h_mean * 1e-3),
        }
        # The haystack lies in HBM as two 16-byte planes (prefix / suffix) and the scan's prefix filter streams the
        # first one only: the bytes the kernel must move are 16 per row ("algorithmic_bytes"), the database it covers
        # is 32 per row ("db_gbs", the metric BASELINE.json names).  With half the bytes to move the kernel is bound
        # by the POPC (XU) pipe: 2 POPC per (row, needle) pair at 16 per clock and SM.
        sm_clk = float(clocks.get("sm_mhz") or 1965.0) * 1e6
        popc_peak = 16.0 * sm_count * sm_clk
        line_extra["scan"] = {
            "kernel": "%s: 8 needles x %d rows, threshold %d" % (
                "match_rows_m_kernel (tcgen05.mma kind::mxf4, database rows in the M dimension, operand A written to tensor "
                "memory from registers)" if rows_in_m else "match_tiled_kernel (POPC prefix filter)", shard_rows, THRESHOLD),
            "us": s_mean * 1e3, "us_median": s_med * 1e3, "us_min": s_min * 1e3,
            "db_gbs": scan_bytes / (s_mean * 1e-3) / 1e9,
            "db_gbs_note": "database bytes covered per second (32 B per row); may exceed the HBM peak because only the prefix plane is read",
            "algorithmic_bytes": shard_rows * 16, "hbm_gbs": shard_rows * 16 / (s_mean * 1e-3) / 1e9,
            "bound": ("alu pipe (row -> e2m1 operand masking + sign test + loop: ~60 instructions per row, ncu: 58 % busy) + ~10 us launch / ramp" if rows_in_m else
                      "xu (POPC pipe) + launch latency"), "peak": hbm_peak,
            "frac": shard_rows * 16 / (s_mean * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src,
            "popc_scan": {
                "kernel": "match_tiled_kernel forced (ALGO_TILED): the XU-bound alternative, ALGO_AUTO's choice below 1M rows",
                "us": p_mean * 1e3, "popc_per_pair": 2, "popc_rate": 2.0 * shard_rows * 8 / (p_mean * 1e-3),
                "popc_peak": popc_peak, "popc_frac": 2.0 * shard_rows * 8 / (p_mean * 1e-3) / popc_peak,
                "popc_peak_source": "16 POPC / clk / SM x %d SMs x %.0f MHz (median SM clock of this run)" % (sm_count, sm_clk / 1e6)},
            "us_l2_resident": None if l2_mean is None else l2_mean * 1e3,
            "us_l2_resident_note": "back-to-back launches without the L2 flush (the shard's 16 B per row prefix plane fits L2); "
                                   "`us` is the flushed number",
            "rows_per_s": shard_rows / (s_mean * 1e-3), "traffic": traffic_scan,
            "traffic_source": TRAFFIC_SOURCE,
            "needle_batches_us": batch_us,
            "sharded_p2p_us": sharded_p2p, "sharded_p2p_us_on_root": sharded_p2p_root,
            "sharded_p2p_host_stamps_us": p2p_stamps,
            "sharded_p2p_forms": None if n_gpus == 1 else p2p_forms,
            "sharded_p2p_form": None if n_gpus == 1 else ("fused" if rows_in_m else "separate"),
            "sharded_p2p_path": None if n_gpus == 1 else
            "8 needles in HOST memory on rank 0 -> written into every rank's mailbox over NVLink (CUDA IPC) -> every rank "
            "waits for its mailbox's flag, scans its shard (L2 evicted before every query), writes its count into rank 0's "
            "mailbox -> rank 0 collects -> host bit mask of the matching shards; no NCCL call.  fused: all of it inside one "
            "launch of the rows-in-M scan per rank; separate: post / wait / scan / publish / collect kernels "
            "(ShardedMatcher.scan_from; median wall clock of 30 queries, max over ranks / on the root)",
            "sharded_e2e_us": sharded_e2e,
            "sharded_e2e_path": None if n_gpus == 1 else
            "8 needles on rank 0 -> NCCL broadcast -> every rank scans its shard -> NCCL all-reduce(MAX) of the flag -> "
            "host bool (median wall clock of 30 queries, max over ranks)",
        }

    # ---- BASELINE configs 5 and 4 ---------------------------------------------------
    c5 = None if args.no_config5 else config5_leg(args, yp, dist, torch, dev, rank, n_gpus, local_rank, frames, mxf4_peak,
                                                   barrier, max_over_ranks)
    c4 = None if args.no_config4 else config4_leg(args, yp, sharded, dist, torch, dev, rank, n_gpus, local_rank, sm,
                                                   barrier, max_over_ranks)

    # ---- end to end through the host-pointer C ABI ---------------------------------
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 3))
        host_frames = torch.empty((B, 512, 512), dtype=torch.float32, pin_memory=True)
        host_frames.copy_(frames)
        host_hashes = np.zeros((B, 8, 32), dtype=np.uint8)
        host_hashes_t = torch.from_numpy(host_hashes)

        def e2e_step():
            if n_gpus == 1:
                # one pipelined C-ABI call: H2D of chunk i+1 overlaps hash + match of chunk i
                return matcher.screen(host_frames.numpy(), hashes=host_hashes)
            # the same pipeline across ranks: chunked H2D || hash -> NCCL all-gather -> match vs the shard,
            # then the gather + merge of the match lists (yume_pdq_b200/sharded.py)
            m = sm.screen(host_frames, hashes_out=outs["hashes"])
            host_hashes_t.copy_(outs["hashes"])
            return m

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            m = e2e_step()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": n_gpus * B * e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(B * 512 * 512 * 4),
               "d2h_bytes_per_step": int(B * 8 * 32 + 16 + len(m) * 16) if n_gpus == 1 else
               int(B * 8 * 32 + 4 + (1 << 16) * 16 * n_gpus),
               "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
               "h2d_gbs_per_rank": B * 512 * 512 * 4 * e2e_steps / dt / 1e9,
               "path": ("yume_pdq_cuda_screen_batch(pinned host frames -> host hashes + sorted host match list)"
                        if n_gpus == 1 else
                        "ShardedMatcher.screen(pinned host frames): chunked H2D || yume_pdq_cuda_hash_batch -> NCCL "
                        "all-gather -> yume_pdq_cuda_matcher_find_device_base, then gather + merge of match lists")}
        del host_frames

        # the same call with 8-bit luma frames (what the reference's `pipe` ingests, main.rs:486-510): the input
        # conversion is fused into the hash kernel, so a frame crosses PCIe as 256 KiB instead of 1 MiB
        dev8 = torch.empty((B, 512, 512), dtype=torch.uint8, device=dev)
        yp.synth_pixels(dev8, B, yp.PIXEL_LUMA8, first_frame=rank * B, seed=SEED, device=local_rank)
        want8 = kern.hash_batch(dev8, want_quality=False, want_thresholds=False, pixel_format=yp.PIXEL_LUMA8)["hashes"]
        host8 = torch.empty((B, 512, 512), dtype=torch.uint8, pin_memory=True)
        host8.copy_(dev8)
        del dev8

        def e2e8_step():
            if n_gpus == 1:
                return matcher.screen(host8.numpy(), hashes=host_hashes, pixel_format=yp.PIXEL_LUMA8)
            m8 = sm.screen(host8, hashes_out=outs["hashes"], pixel_format=yp.PIXEL_LUMA8)
            host_hashes_t.copy_(outs["hashes"])
            return m8

        e2e8_step()
        if not np.array_equal(host_hashes, want8.cpu().numpy()):
            raise SystemExit("bench.py: LUMA8 end-to-end hashes differ from the device-resident path")
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            m8 = e2e8_step()
        barrier()
        dt8 = max_over_ranks(time.perf_counter() - t0)
        e2e["luma8"] = {"value": n_gpus * B * e2e_steps / dt8, "unit": UNIT,
                        "h2d_bytes_per_step": int(B * 512 * 512), "ms_per_step": dt8 / e2e_steps * 1e3,
                        "d2h_bytes_per_step": int(B * 8 * 32 + 16 + len(m8) * 16) if n_gpus == 1 else
                        int(B * 8 * 32 + 4 + (1 << 16) * 16 * n_gpus),
                        "path": "the same call with pinned 8-bit luma frames (yume_pdq_cuda_screen_batch_pixels, "
                                "YUME_PDQ_CUDA_PIXEL_LUMA8): conversion fused into the hash kernel"}
        del host8

        # the reference's single-frame C ABI (what integration/hash.py calls): per-call latency through ctypes with
        # preallocated buffers, as README.md:170-173 measures it
        if rank == 0:
            import ctypes
            one = frames[0].cpu().numpy()
            one_pinned = torch.empty((512, 512), dtype=torch.float32).pin_memory()
            one_pinned.copy_(frames[0])
            torch.cuda.synchronize()
            o_thr = ctypes.c_float(0.0)
            o_hash = np.zeros(32, dtype=np.uint8)
            o_buf1 = np.zeros((128, 128), dtype=np.float32)
            o_tmp = np.zeros(128, dtype=np.float32)
            o_pdqf = np.zeros((16, 16), dtype=np.float32)
            fn = yp.lib.yume_pdq_hash_smart_kernel
            outs = (ctypes.byref(o_thr), ctypes.c_void_p(o_hash.ctypes.data), ctypes.c_void_p(o_buf1.ctypes.data),
                    ctypes.c_void_p(o_tmp.ctypes.data), ctypes.c_void_p(o_pdqf.ctypes.data))

            def per_call(ptr):
                src = ctypes.c_void_p(ptr)
                for _ in range(20):
                    fn(src, *outs)
                ts = []
                for _ in range(300):
                    t0 = time.perf_counter()
                    fn(src, *outs)
                    ts.append(time.perf_counter() - t0)
                return float(np.median(ts)) * 1e6, float(np.mean(ts)) * 1e6

            med_pg, mean_pg = per_call(one.ctypes.data)
            ref_hash = o_hash.copy()
            med_pin, mean_pin = per_call(one_pinned.data_ptr())
            assert (o_hash == ref_hash).all()
            med_dev, mean_dev = per_call(frames[0].data_ptr())
            assert (o_hash == ref_hash).all()
            e2e["single_frame_abi"] = {
                "us_per_call": med_pg, "us_per_call_mean": mean_pg,
                "us_per_call_pinned_input": med_pin, "us_per_call_device_input": med_dev,
                "path": "yume_pdq_hash_smart_kernel through ctypes with preallocated buffers, median of 300 calls.  Frame in "
                        "PAGEABLE host memory: four quarter-frame copies into pinned memory (non-temporal stores), each handed "
                        "to the copy engine as soon as it is there, then ONE launch (hash_one_kernel: 128 CTAs for the tent "
                        "rows, the last one to finish runs dct / median / gradients / hash bits); results and the completion "
                        "flag are written to mapped host memory and polled -- no D2H copy, no stream sync.  pinned_input: one "
                        "DMA; device_input: frame already in HBM",
                "reference_published_us": 81.0, "reference_source": "README.md:170-173 (AVX2 build through ctypes, Ryzen 9 7950X)"}

    cpu_bl = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        cpu_bl = cpu_baseline_leg(args)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (hash) / e2m1 x e2m1 -> f32 exact (match, tensor cores) / u32 popcount (8-needle scan)",
            "data": "synthetic", "config": workload_config(n_gpus),
            "e2e": e2e, "gpu_launches": int(lt.item()), "clocks": clocks,
            "matches_per_step": total_matches, "planted": N_PLANT * n_gpus,
            "parity_checked": bool(parity and parity.get("checked")), "parity": parity,
            "config5": c5, "config4": c4, "worst_case": worst, "step_timeline": timeline,
            "needle_gather": None if n_gpus == 1 else
            ("NCCL all-gather + two packing copies" if exchange is None else
             "one push kernel over peer memory before the matcher (yume_pdq_cuda_exchange_*: CUDA IPC buffers, NVLink stores, "
             "flag words)"),
            "cpu_baseline": cpu_bl,
        }
        line.update(line_extra)
        emit(line)
    db.close()
    if n_gpus > 1:
        dist.barrier()
        dist.destroy_process_group()


def worst_case_leg(args, yp, torch, dev, local_rank, rows_n, flush):
    """ALGO_AUTO on haystacks that are NOT random (one GPU, this rank's row count), threshold 31, 32768 needles:
      * "families": 0.1 % of the rows lie within 0..40 bits of one of 32 base hashes and the needles are 32 'videos' of
        1024 near-identical frames (within 6 bits of a base) -- every such row is a prefix-filter candidate against 1024
        needles (3e-5 of all pairs), most of them real matches.  All needles of a tile flag the same accumulator column,
        so few accumulator batches take the filter's rare path and the filter stays the fast kernel;
      * "scattered": 10 % of the rows are each within 0..40 bits of ONE random needle -- a million separate flagged
        batches, most of them matches, which the unfiltered kernel pays ~115 ns each for: the filter stays;
      * "prefix_twins": 25 % of the rows share their first 128 bits (within 0..31 bits) with one random needle but not
        the rest -- candidates that are not matches, 4 % of the accumulator batches flag: ALGO_AUTO falls back to the
        unfiltered kernel.
    For each: steady-state rate of ALGO_AUTO, of the prefix filter forced on, of the unfiltered kernel forced on.
    `worst_case.tpairs_per_s` is the slowest of the ALGO_AUTO rates."""
    rng = np.random.default_rng(4242)
    q = 32768

    def flipped(base_rows, k):
        """each 32-byte row with exactly k[i] distinct random bits flipped"""
        order = np.argsort(rng.random((base_rows.shape[0], 256)), axis=1)
        mask = (order < k[:, None]).astype(np.uint8)
        return np.packbits(np.unpackbits(base_rows, axis=1) ^ mask, axis=1)

    def dist(a, b):
        return np.bitwise_count(a ^ b).sum(axis=-1, dtype=np.int32)

    def scenario(kind):
        if kind == "families":
            n_cl, per = 32, 1024
            bases = rng.integers(0, 256, size=(n_cl, 32), dtype=np.uint8)
            needles = flipped(np.repeat(bases, per, axis=0), rng.integers(0, 7, size=q))
            n_near = max(n_cl, rows_n // 1000)
            near_cl = rng.integers(0, n_cl, size=n_near)
            near = flipped(bases[near_cl], rng.integers(0, 41, size=n_near))
            expect = 0
            for c in range(n_cl):
                rc = near[near_cl == c]
                if rc.shape[0]:
                    expect += int((dist(rc[:, None, :], needles[None, c * per:(c + 1) * per, :]) <= THRESHOLD).sum())
            what = "%d rows (0.1 %%) within 0..40 bits of one of %d bases; needles = %d x %d frames within 6 bits of a base" % (
                n_near, n_cl, n_cl, per)
            cand_frac = n_near * per / (rows_n * q)
        elif kind == "scattered":
            needles = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
            n_near = rows_n // 10
            of = rng.integers(0, q, size=n_near)
            near = np.concatenate([flipped(needles[of[i:i + 65536]], rng.integers(0, 41, size=len(of[i:i + 65536])))
                                   for i in range(0, n_near, 65536)])
            expect = int((dist(near, needles[of]) <= THRESHOLD).sum())
            what = "%d rows (10 %%) each within 0..40 bits of one random needle of %d" % (n_near, q)
            cand_frac = n_near / (rows_n * q)
        else:
            needles = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
            n_near = rows_n // 4
            of = rng.integers(0, q, size=n_near)
            near = rng.integers(0, 256, size=(n_near, 32), dtype=np.uint8)          # random suffix ...
            for i in range(0, n_near, 65536):                                        # ... prefix of a needle, 0..31 bits flipped
                blk = of[i:i + 65536]
                pre = np.concatenate([needles[blk, :16], np.zeros((len(blk), 16), dtype=np.uint8)], axis=1)
                order = np.argsort(rng.random((len(blk), 128)), axis=1)
                mask = np.concatenate([(order < rng.integers(0, 32, size=len(blk))[:, None]).astype(np.uint8),
                                       np.zeros((len(blk), 128), dtype=np.uint8)], axis=1)
                near[i:i + len(blk), :16] = np.packbits(np.unpackbits(pre, axis=1) ^ mask, axis=1)[:, :16]
            expect = int((dist(near, needles[of]) <= THRESHOLD).sum())
            what = "%d rows (25 %%) whose first 128 bits are within 0..31 bits of one random needle of %d, the rest random" % (n_near, q)
            cand_frac = n_near / (rows_n * q)
        where = rng.choice(rows_n, size=n_near, replace=False)
        rows_t = torch.empty((rows_n, 32), dtype=torch.uint8, device=dev)
        yp.synth_rows(rows_t, rows_n, first_row=1 << 40, seed=SEED, device=local_rank)
        rows_t[torch.from_numpy(where).to(dev)] = torch.from_numpy(near).to(dev)
        dbc = yp.CudaVectorDatabase(rows_t, device=local_rank)
        del rows_t
        mc = yp.CudaMatcher(dbc, THRESHOLD)
        cap = max(1 << 16, 1 << int(np.ceil(np.log2(expect + 4096))))
        out = torch.zeros((cap, 16), dtype=torch.uint8, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        nd = torch.from_numpy(needles).to(dev)

        def run(algo, reps):
            ts = []
            for _ in range(reps):
                cnt.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                mc.find_device(nd, q, out, cap, cnt, algo=algo)
                b.record()
                torch.cuda.synchronize()       # the host reads the previous launch's statistics before the next one
                ts.append(a.elapsed_time(b))
                if int(cnt.item()) != expect:
                    raise SystemExit(f"bench.py: {kind} haystack, algo {algo}: {int(cnt.item())} matches, numpy expects {expect}")
            return ts

        first = run(yp.ALGO_AUTO, 1)[0]                  # a fresh haystack: the prefix filter runs once and reports
        run(yp.ALGO_AUTO, 2)
        auto = run(yp.ALGO_AUTO, 24)                      # steady state, incl. the probes of the filter while in fallback
        fallback, interval = dbc.auto_state
        full = run(yp.ALGO_TENSOR_FULL, 4)
        paired = run(yp.ALGO_TENSOR_PAIRED, 3)
        pairs = rows_n * q
        tp = lambda ms: pairs / (ms * 1e-3) / 1e12
        res = {
            "workload": what, "candidate_pairs_fraction": cand_frac, "matches": expect, "matches_checked_against": "numpy popcounts",
            "auto_ms_mean": float(np.mean(auto)), "auto_ms_max": float(np.max(auto)), "auto_tpairs_per_s": tp(float(np.mean(auto))),
            "auto_in_fallback": bool(fallback), "auto_probe_interval": interval, "first_launch_ms": first,
            "unfiltered_ms": float(np.mean(full)), "unfiltered_tpairs_per_s": tp(float(np.mean(full))),
            "prefix_filter_forced_ms": float(np.mean(paired)), "prefix_filter_forced_tpairs_per_s": tp(float(np.mean(paired))),
            "auto_vs_best_forced": min(float(np.mean(full)), float(np.mean(paired))) / float(np.mean(auto)),
        }
        dbc.close()
        del out, nd
        return res

    fam, sca, twi = scenario("families"), scenario("scattered"), scenario("prefix_twins")
    return {"rows": rows_n, "needles": q, "threshold": THRESHOLD, "families": fam, "scattered": sca, "prefix_twins": twi,
            "tpairs_per_s": min(fam["auto_tpairs_per_s"], sca["auto_tpairs_per_s"], twi["auto_tpairs_per_s"]),
            "fallback_rule": "15 ns x flagged accumulator batches - 115 ns x matches > 3.3e-5 ns x pairs and flagged > 3 % of "
                             "the batches (MatchStats, yume_pdq_b200/csrc/common.cuh)"}


PARITY_ROWS = 200_000
PARITY_SCALAR_ROWS = 16_384      # rows (every stride-th of the 200,000) the scalar oracle checks, divided by N


def parity_leg(args, yp, sharded, dist, torch, dev, rank, n_gpus, local_rank, local_needles, needles_all, q_all,
               exchange_find_all=None):
    """Full equality of the sharded matcher's merged match list with the CPU oracle, on a sub-sampled problem: a
    200,000-row synthetic DB (row-range sharded over the N ranks exactly like the timed one) x ALL q_all needles of
    the job.  Rank 0 compares (row, needle, distance) lists: against the AVX-512 port of the reference matcher on all
    rows, and against the scalar oracle (oracle/pdq_oracle.c, the pinned one) on every stride-th row, where half of
    the plants sit.  A mismatch aborts the bench."""
    if n_gpus > 1:
        dist.all_gather_into_tensor(needles_all, local_needles)
    torch.cuda.synchronize()
    needles_host = needles_all.cpu().numpy()
    seed = SEED + 1
    smp = sharded.ShardedMatcher(
        PARITY_ROWS, THRESHOLD,
        lambda b, e, d: yp.CudaVectorDatabase(device=d, global_row_offset=b, synthetic_rows=e - b, seed=seed),
        device=local_rank)
    n_scalar = max(1024, PARITY_SCALAR_ROWS // n_gpus)
    stride = PARITY_ROWS // n_scalar
    rng = np.random.default_rng(99)            # the same plants on every rank

    def write(row, bits32):
        smp.db.write_rows(row - smp.begin, bits32[None])

    planted = plant_rows(rng, needles_host, 96, 0, PARITY_ROWS, write, smp.begin, smp.end, stride=stride)
    planted += plant_rows(rng, needles_host, 96, 0, PARITY_ROWS, write, smp.begin, smp.end, stride=1)
    t0 = time.perf_counter()
    got = smp.find_all(local_needles, cap=1 << 14)        # all-gather needles -> shard scans -> gather + merge
    t_gpu = time.perf_counter() - t0
    # the timed step's own data path (peer exchange + split match) on the same problem must give the same list
    got_x = exchange_find_all(smp.matcher) if exchange_find_all is not None else None
    smp.close()
    if got_x is not None:
        kx = [(int(r), int(n), int(d)) for r, n, d in zip(got_x["row"], got_x["needle"], got_x["dist"])]
        kg = [(int(r), int(n), int(d)) for r, n, d in zip(got["row"], got["needle"], got["dist"])]
        if kx != kg:
            raise SystemExit(f"bench.py: PARITY FAILURE: rank {rank}: peer-exchange path ({len(kx)} matches) != "
                             f"ShardedMatcher.find_all ({len(kg)} matches)")
    out = {"checked": False, "exchange_path_checked": got_x is not None, "rows": PARITY_ROWS, "needles": int(q_all), "threshold": THRESHOLD,
           "matches": int(len(got)), "planted": len(planted), "sharded_find_all_ms": t_gpu * 1e3}
    if rank != 0:
        return out
    from oracle import pyoracle
    pyoracle.build()
    orc = pyoracle.get()
    rows = orc.synth_db(PARITY_ROWS, 0, seed)
    rng = np.random.default_rng(99)
    plant_rows(rng, needles_host, 96, 0, PARITY_ROWS, lambda r, b: rows.__setitem__(r, b), stride=stride)
    plant_rows(rng, needles_host, 96, 0, PARITY_ROWS, lambda r, b: rows.__setitem__(r, b), stride=1)
    key = lambda m: [(int(r), int(n), int(d)) for r, n, d in zip(m["row"], m["needle"], m["dist"])]
    got_k = key(got)
    cpu = pyoracle.CpuBaseline()
    t0 = time.perf_counter()
    want_port = key(cpu.match(rows, needles_host, THRESHOLD, True, cpu.hw_threads))
    t_port = time.perf_counter() - t0
    if got_k != want_port:
        raise SystemExit("bench.py: PARITY FAILURE: sharded match list != AVX-512 port of the reference matcher "
                         f"({len(got_k)} vs {len(want_port)} matches)")
    t0 = time.perf_counter()
    sub = np.ascontiguousarray(rows[::stride][:n_scalar])
    want_scalar = [(r * stride, n, d) for r, n, d in key(orc.matcher_find(sub, needles_host, THRESHOLD))]
    t_scalar = time.perf_counter() - t0
    got_sub = [t for t in got_k if t[0] % stride == 0 and t[0] // stride < n_scalar]
    if got_sub != want_scalar:
        raise SystemExit("bench.py: PARITY FAILURE: sharded match list != scalar oracle on the sub-sampled rows "
                         f"({len(got_sub)} vs {len(want_scalar)} matches)")
    expect = sum(1 for p in planted if p[2] <= THRESHOLD)
    out.update(checked=True, expected_from_plants_at_least=expect,
               oracle=("full equality of (row, needle, dist): AVX-512 port of CpuMatcher::find on all %d rows (%.1f s, %d "
                       "threads) and scalar oracle on every %d-th row (%d rows, %d matches, %.1f s)"
                       % (PARITY_ROWS, t_port, cpu.hw_threads, stride, n_scalar, len(want_scalar), t_scalar)))
    return out


def config5_leg(args, yp, dist, torch, dev, rank, n_gpus, local_rank, frames, mxf4_peak, barrier, max_over_ranks):
    """BASELINE config 5 (INT/tensor regime): a 1e9-row DB over 8 GPUs = 125M synthetic rows (4 GB) per GPU -- at N < 8
    the DB is proportionally smaller (125M x N rows), so per-GPU work is constant --, batches of 1024 images x 8
    dihedrals: each rank hashes 1024 / N resident frames, the 8192 needles are all-gathered and matched against the
    rank's shard in ONE launch."""
    rows_per_gpu = args.c5_rows_per_gpu
    imgs = 1024 // n_gpus
    q_loc, q_all = imgs * 8, 1024 * 8
    kern = yp.CudaKernel(local_rank)
    db5 = yp.CudaVectorDatabase(device=local_rank, global_row_offset=rank * rows_per_gpu, synthetic_rows=rows_per_gpu,
                                seed=SEED + 5)
    m5 = yp.CudaMatcher(db5, THRESHOLD)
    fr = frames[:imgs]
    hashes = torch.empty((imgs, 8, 32), dtype=torch.uint8, device=dev)
    needles = torch.empty((q_all, 32), dtype=torch.uint8, device=dev) if n_gpus > 1 else hashes.reshape(-1, 32)
    CAP5 = 1 << 14
    out = torch.zeros((CAP5, 16), dtype=torch.uint8, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    kern.hash_batch(fr, want_quality=False, want_thresholds=False, out={"hashes": hashes})
    torch.cuda.synchronize()
    hh = hashes.reshape(-1, 32).cpu().numpy()
    rng = np.random.default_rng(555 + rank)
    planted = plant_rows(rng, hh, 16, 0, rows_per_gpu, lambda r, b: db5.write_rows(r, b[None]))
    planted = {(rank * rows_per_gpu + r, rank * q_loc + j) for r, j, d in planted if d <= THRESHOLD}

    def hash_():
        kern.hash_batch(fr, want_quality=False, want_thresholds=False, out={"hashes": hashes})

    def gather_():
        if n_gpus > 1:
            dist.all_gather_into_tensor(needles, hashes.reshape(-1, 32))

    def match_():
        cnt.zero_()
        m5.find_device(needles, q_all, out, CAP5, cnt)

    def step():
        hash_()
        gather_()
        match_()

    def timed(fn, reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b) / reps)

    for _ in range(3):
        step()
    steps = 5
    ms = timed(step, steps)
    ms_hash, ms_gather, ms_match = timed(hash_, 5), timed(gather_, 5), timed(match_, 3)
    n = int(cnt.item())
    got = out[:min(n, CAP5)].cpu().numpy().view(yp.MATCH_DTYPE).reshape(-1)
    got_set = {(int(r), int(j)) for r, j in zip(got["row"], got["needle"])}
    if not planted <= got_set:
        raise SystemExit(f"bench.py: config5: rank {rank}: planted matches missing")
    # the headline 8-needle scan on this 125M-row shard (2 GB prefix plane, far beyond L2): ALGO_AUTO (rows-in-M tensor
    # kernel) and the POPC kernel must report the same matches; the first 8 needles hold planted ones on some ranks
    n8 = needles[:8].contiguous()
    scan_ms, scan_cnt = {}, {}
    for name, algo in (("auto", yp.ALGO_AUTO), ("popc", yp.ALGO_TILED)):
        scan_ms[name] = timed(lambda: (cnt.zero_(), m5.find_device(n8, 8, out, CAP5, cnt, algo=algo)), 5)
        scan_cnt[name] = int(cnt.item())
    if scan_cnt["auto"] != scan_cnt["popc"]:
        raise SystemExit(f"bench.py: config5: rank {rank}: 8-needle scan: {scan_cnt['auto']} vs {scan_cnt['popc']} matches")
    db5.close()
    pairs_gpu = rows_per_gpu * q_all
    ach = pairs_gpu * 256 / (ms_match * 1e-3) / 1e12
    return {
        "workload": "%d synthetic rows per GPU (%d total, %.1f GB per shard), 1024 images x 8 dihedrals per batch, ONE "
                    "%d-needle launch per batch, threshold %d" % (rows_per_gpu, rows_per_gpu * n_gpus, rows_per_gpu * 32 / 1e9,
                                                                  q_all, THRESHOLD),
        "rows_per_gpu": rows_per_gpu, "db_rows_total": rows_per_gpu * n_gpus, "needles_per_batch": q_all,
        "ms_per_batch": ms, "steps": steps, "batches_per_s": 1e3 / ms, "images_per_s": 1024 / (ms * 1e-3),
        "tpairs_per_s_total": pairs_gpu * n_gpus / (ms * 1e-3) / 1e12, "tpairs_per_s_per_gpu": pairs_gpu / (ms * 1e-3) / 1e12,
        "split_ms": {"hash": ms_hash, "all_gather": ms_gather, "match": ms_match},
        "launches_per_batch": "1 hash + 1 needle expansion + 1 match + 1 candidate verification",
        "roofline": {"bound": "tensor", "achieved": ach, "peak": mxf4_peak, "unit": "TFLOP/s", "frac": ach / mxf4_peak,
                     "peak_is_self_measured": True, "tensor_ops_per_pair": 256, "ms_per_launch": ms_match,
                     "share_of_step": ms_match / ms},
        "matches_per_batch": n, "planted_found": len(planted), "scaling": "weak (rows per GPU fixed)",
        "scan_8_needles": {"us": scan_ms["auto"] * 1e3, "us_popc_kernel": scan_ms["popc"] * 1e3,
                           "db_gbs": rows_per_gpu * 32 / (scan_ms["auto"] * 1e-3) / 1e9,
                           "hbm_gbs": rows_per_gpu * 16 / (scan_ms["auto"] * 1e-3) / 1e9,
                           "db_gbs_total": n_gpus * rows_per_gpu * 32 / (scan_ms["auto"] * 1e-3) / 1e9,
                           "note": "every rank scans its 125M-row shard with the same 8 needles (max over ranks); db_gbs = "
                                   "database bytes covered per second (32 B per row), hbm_gbs = bytes read (16 B per row)"},
    }


def config4_leg(args, yp, sharded, dist, torch, dev, rank, n_gpus, local_rank, sm, barrier, max_over_ranks):
    """BASELINE config 4 (video screening): 65,536 device-generated 512x512 f32 frames in total, hashed and matched
    against the 10M-row DB sharded over the N ranks, through ShardedMatcher.screen (chunked hash -> NCCL all-gather of
    the chunk's needles -> match against the shard, then gather + merge of the match lists)."""
    total = args.c4_frames
    F = total // n_gpus
    frames4 = torch.empty((F, 512, 512), dtype=torch.float32, device=dev)
    yp.synth_frames(frames4, F, first_frame=10_000_000 + rank * F, seed=SEED, kind=1, device=local_rank)
    hashes4 = torch.empty((F, 8, 32), dtype=torch.uint8, device=dev)
    chunk = 2048
    kern = yp.CudaKernel(local_rank)
    kern.hash_batch(frames4, want_quality=False, want_thresholds=False, out={"hashes": hashes4})
    torch.cuda.synchronize()
    hh = hashes4.reshape(-1, 32).cpu().numpy()
    rng = np.random.default_rng(444 + rank)
    shard_rows = sm.end - sm.begin
    planted = plant_rows(rng, hh, 16, 0, shard_rows, lambda r, b: sm.db.write_rows(r, b[None]))
    planted = {(sm.begin + r, rank * F * 8 + j) for r, j, d in planted if d <= THRESHOLD}

    def run():
        return sm.screen(frames4, hashes_out=hashes4, cap=1 << 14, chunk=chunk)

    run()
    barrier()
    reps = 2
    t0 = time.perf_counter()
    for _ in range(reps):
        m = run()
    barrier()
    dt = max_over_ranks((time.perf_counter() - t0) / reps)
    got_set = {(int(r), int(j)) for r, j in zip(m["row"], m["needle"])}
    if not planted <= got_set:
        raise SystemExit(f"bench.py: config4: rank {rank}: planted matches missing")

    def timed(fn, reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b) / reps)

    ms_hash = timed(lambda: kern.hash_batch(frames4, want_quality=False, want_thresholds=False, out={"hashes": hashes4}), 2)
    all_needles = torch.empty((total * 8, 32), dtype=torch.uint8, device=dev) if n_gpus > 1 else hashes4.reshape(-1, 32)
    ms_gather = timed(lambda: dist.all_gather_into_tensor(all_needles, hashes4.reshape(-1, 32)), 2) if n_gpus > 1 else 0.0
    out = torch.zeros((1 << 14, 16), dtype=torch.uint8, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    ms_match = timed(lambda: (cnt.zero_(), sm.matcher.find_device(all_needles, total * 8, out, 1 << 14, cnt)), 2)
    del frames4
    pairs_gpu = shard_rows * total * 8
    return {
        "workload": "%d device-generated 512x512 f32 frames in total (%d per GPU), 8 dihedral hashes each, matched vs the "
                    "%d-row DB sharded over %d GPU(s), threshold %d" % (total, F, DB_ROWS, n_gpus, THRESHOLD),
        "frames_total": total, "frames_per_gpu": F, "needles_total": total * 8, "chunk_frames": chunk if n_gpus > 1 else F,
        "ms_per_pass": dt * 1e3, "images_per_s": total / dt, "passes": reps,
        "tpairs_per_s_total": pairs_gpu * n_gpus / dt / 1e12,
        "path": "ShardedMatcher.screen (wall clock incl. the host-side gather + merge of the match lists)",
        "split_ms": {"hash_all_frames": ms_hash, "all_gather_all_needles": ms_gather, "match_all_needles_one_launch": ms_match,
                     "remainder_pipeline_gather_merge": dt * 1e3 - ms_hash - ms_gather - ms_match},
        "matches": int(len(m)), "planted_found": len(planted),
    }


if __name__ == "__main__":
    main()
