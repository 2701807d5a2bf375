"""Developer micro-benchmarks (not the contract bench): CUDA-event timings of the individual kernels."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yume_pdq_b200 as yp


def timeit(fn, iters=20, warm=3, flush=None):
    """median / best in microseconds; everything is enqueued asynchronously so no host launch gap is timed"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    if flush is not None:
        flush.sum()
    for e0, e1 in ev:
        if flush is not None:
            flush.sum()      # READ 256 MB: evicts L2 without leaving dirty lines to write back during the timed kernel
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    ts = sorted(e0.elapsed_time(e1) * 1e3 for e0, e1 in ev)
    return ts[len(ts) // 2], ts[0]


def main():
    what = sys.argv[1:] or ["scan", "hash", "tiled"]
    torch.cuda.init()
    print(torch.cuda.get_device_name(0))
    flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")   # 256 MB
    rng = np.random.default_rng(0)
    if "scan" in what:
        n = (10_000_000 // 2048 + 1) * 2048
        db = yp.CudaVectorDatabase(synthetic_rows=n)
        m = yp.CudaMatcher(db, 31)
        out = torch.zeros((1 << 16, 16), dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        for q in (8, 16, 32):
            needles = torch.from_numpy(rng.integers(0, 256, size=(q, 32), dtype=np.uint8)).cuda()
            for name, algo in (("exact", yp.ALGO_EXACT), ("tiled", yp.ALGO_TILED)):
                med, best = timeit(lambda: m.find_device(needles, q, out, 1 << 16, cnt, algo=algo), flush=flush)
                print(f"scan 10M x {q:3d} {name:7s}: median {med:8.1f} us  best {best:8.1f} us  "
                      f"{n * 32 / med / 1e3:7.1f} GB/s  {n * q / med / 1e6:8.3f} Tpairs/s  hits={int(cnt.item())}")
                cnt.zero_()
        db.close()
    if "tc4v" in what:   # packed-FP4 tensor matcher variants (YPDQ_TC4_VARIANT, one process per variant)
        n = (10_000_000 // 2048 + 1) * 2048
        db = yp.CudaVectorDatabase(synthetic_rows=n)
        m = yp.CudaMatcher(db, 31)
        out = torch.zeros((1 << 20, 16), dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        for q in (1024, 4096, 32768):
            needles = torch.from_numpy(rng.integers(0, 256, size=(q, 32), dtype=np.uint8)).cuda()
            med, best = timeit(lambda: m.find_device(needles, q, out, 1 << 20, cnt, algo=yp.ALGO_TENSOR), iters=5, warm=1)
            print(f"tc4 variant {os.environ.get('YPDQ_TC4_VARIANT', '-')}: match 10M x {q:6d}: median {med / 1e3:9.2f} ms  "
                  f"{n * q / med / 1e6:8.3f} Tpairs/s  hits={int(cnt.item())}")
            cnt.zero_()
        db.close()
    if "tc4p" in what:   # full 256-bit vs 128-bit prefix tensor matcher (run once per YUME_PDQ_CUDA_LIB variant)
        tag = os.path.basename(os.environ.get("YUME_PDQ_CUDA_LIB", "default"))
        peak, clk = yp.measure_mxf4_peak(0)
        print(f"[{tag}] mxf4 issue rate: {peak:.0f} Tops/s, {clk:.1f} SM clk per MMA (M128 N240 K64)")
        n = (10_000_000 // 2048 + 1) * 2048
        db = yp.CudaVectorDatabase(synthetic_rows=n)
        m = yp.CudaMatcher(db, 31)
        out = torch.zeros((1 << 20, 16), dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        for q in (512, 1024, 4096, 32768):
            needles = torch.from_numpy(rng.integers(0, 256, size=(q, 32), dtype=np.uint8)).cuda()
            for name, algo in (("full", yp.ALGO_TENSOR_FULL), ("prefix", yp.ALGO_TENSOR_PREFIX), ("paired", yp.ALGO_TENSOR_PAIRED)):
                med, best = timeit(lambda: m.find_device(needles, q, out, 1 << 20, cnt, algo=algo), iters=5, warm=1)
                print(f"[{tag}] match 10M x {q:6d} {name:6s}: median {med / 1e3:9.2f} ms  {n * q / med / 1e6:8.3f} Tpairs/s  "
                      f"hits={int(cnt.item())}")
                cnt.zero_()
        for thr in (36, 40):   # candidate-rate sensitivity of the prefix filter
            mt = yp.CudaMatcher(db, thr)
            needles = torch.from_numpy(rng.integers(0, 256, size=(32768, 32), dtype=np.uint8)).cuda()
            for name, algo in (("prefix", yp.ALGO_TENSOR_PREFIX), ("paired", yp.ALGO_TENSOR_PAIRED)):
                med, best = timeit(lambda: mt.find_device(needles, 32768, out, 1 << 20, cnt, algo=algo), iters=3, warm=1)
                print(f"[{tag}] match 10M x  32768 {name} thr={thr}: median {med / 1e3:9.2f} ms  {n * 32768 / med / 1e6:8.3f} Tpairs/s")
                cnt.zero_()
        db.close()
    if "x2" in what:   # paired tensor matcher only (developer variants: YUME_PDQ_CUDA_LIB=...)
        tag = os.path.basename(os.environ.get("YUME_PDQ_CUDA_LIB", "default"))
        n = (10_000_000 // 2048 + 1) * 2048
        db = yp.CudaVectorDatabase(synthetic_rows=n)
        m = yp.CudaMatcher(db, 31)
        out = torch.zeros((1 << 20, 16), dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        for q in (768, 1024, 1536, 2048, 4096, 32768):
            needles = torch.from_numpy(rng.integers(0, 256, size=(q, 32), dtype=np.uint8)).cuda()
            med, best = timeit(lambda: m.find_device(needles, q, out, 1 << 20, cnt, algo=yp.ALGO_TENSOR_PAIRED), iters=5, warm=1)
            print(f"[{tag}] match 10M x {q:6d} paired: median {med / 1e3:9.2f} ms  {n * q / med / 1e6:8.3f} Tpairs/s  "
                  f"hits={int(cnt.item())}")
            cnt.zero_()
        for thr in (33, 36, 40):
            mt = yp.CudaMatcher(db, thr)
            med, best = timeit(lambda: mt.find_device(needles, 32768, out, 1 << 20, cnt, algo=yp.ALGO_TENSOR_PAIRED), iters=3, warm=1)
            print(f"[{tag}] match 10M x  32768 paired thr={thr}: median {med / 1e3:9.2f} ms  {n * 32768 / med / 1e6:8.3f} Tpairs/s")
            cnt.zero_()
        db.close()
    if "shard" in what:   # the matcher on the row shards of a 10M-row DB over 8 / 4 / 2 GPUs, needle batches of the N-GPU step
        tag = os.path.basename(os.environ.get("YUME_PDQ_CUDA_LIB", "default"))
        out = torch.zeros((1 << 20, 16), dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        for n, qs in ((1_250_048, (32768, 7 * 32768)), (2_500_096, (32768, 3 * 32768)), (10_000_384, (32768,))):
            db = yp.CudaVectorDatabase(synthetic_rows=n)
            m = yp.CudaMatcher(db, 31)
            for q in qs:
                needles = torch.from_numpy(rng.integers(0, 256, size=(q, 32), dtype=np.uint8)).cuda()
                med, best = timeit(lambda: m.find_device(needles, q, out, 1 << 20, cnt), iters=5, warm=1)
                print(f"[{tag}] match {n:9d} x {q:6d}: median {med / 1e3:9.3f} ms  {n * q / med / 1e6:8.3f} Tpairs/s  hits={int(cnt.item())}")
                cnt.zero_()
            db.close()
    if "x1" in what:   # rows-in-M tensor scan only (developer variants: YUME_PDQ_CUDA_LIB=...)
        tag = os.path.basename(os.environ.get("YUME_PDQ_CUDA_LIB", "default"))
        out = torch.zeros((1 << 16, 16), dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        for n in (1_250_048, (10_000_000 // 2048 + 1) * 2048):
            db = yp.CudaVectorDatabase(synthetic_rows=n)
            m = yp.CudaMatcher(db, 31)
            for q in (8, 16, 32, 64):
                needles = torch.from_numpy(rng.integers(0, 256, size=(q, 32), dtype=np.uint8)).cuda()
                med, best = timeit(lambda: m.find_device(needles, q, out, 1 << 16, cnt, algo=yp.ALGO_TENSOR_ROWS), iters=20, warm=3, flush=flush)
                print(f"[{tag}] rows-in-M {n:9d} x {q:3d}: median {med:8.1f} us  best {best:8.1f} us  {n * 16 / med / 1e3:7.1f} GB/s (prefix plane)")
            db.close()
    if "cross" in what:   # crossover between the tiled POPC kernel and the paired tensor kernel
        n = (10_000_000 // 2048 + 1) * 2048
        db = yp.CudaVectorDatabase(synthetic_rows=n)
        m = yp.CudaMatcher(db, 31)
        out = torch.zeros((1 << 20, 16), dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        small = yp.CudaVectorDatabase(synthetic_rows=1_000_000)
        ms = yp.CudaMatcher(small, 31)
        for q in (128, 1024):
            needles = torch.from_numpy(rng.integers(0, 256, size=(q, 32), dtype=np.uint8)).cuda()
            med, best = timeit(lambda: ms.find_device(needles, q, out, 1 << 20, cnt, algo=yp.ALGO_TENSOR_PAIRED), iters=10, warm=2)
            print(f"cross 1M x {q:4d} paired: median {med:9.1f} us")
            cnt.zero_()
        small.close()
        for q in (8, 16, 32, 64, 128, 256, 512, 1024, 2048):
            needles = torch.from_numpy(rng.integers(0, 256, size=(q, 32), dtype=np.uint8)).cuda()
            for name, algo in (("tiled", yp.ALGO_TILED), ("paired", yp.ALGO_TENSOR_PAIRED), ("rows", yp.ALGO_TENSOR_ROWS),
                               ("auto", yp.ALGO_AUTO)):
                if (name == "tiled" and q > 256) or (name == "rows" and q > 32):
                    continue
                med, best = timeit(lambda: m.find_device(needles, q, out, 1 << 20, cnt, algo=algo), iters=10, warm=2, flush=flush)
                print(f"cross 10M x {q:4d} {name:6s}: median {med:9.1f} us  best {best:9.1f} us  {n * q / med / 1e6:8.3f} Tpairs/s  "
                      f"{n * 32 / med / 1e3:7.1f} GB/s")
                cnt.zero_()
        db.close()
    if "fixed" in what:   # fixed cost of a launch: tiny shards (one 960-row unit per SM and multiples)
        out = torch.zeros((1 << 20, 16), dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        for n in (960 * 148, 960 * 148 * 4, 960 * 148 * 16, 1_250_048, 10_000_384):
            db = yp.CudaVectorDatabase(synthetic_rows=n)
            m = yp.CudaMatcher(db, 31)
            for q in (8, 32, 128, 1024):
                needles = torch.from_numpy(rng.integers(0, 256, size=(q, 32), dtype=np.uint8)).cuda()
                for name, algo in (("tiled", yp.ALGO_TILED), ("paired", yp.ALGO_TENSOR_PAIRED), ("rows", yp.ALGO_TENSOR_ROWS)):
                    if (name == "tiled" and q > 128) or (name == "rows" and q > 32):
                        continue
                    med, best = timeit(lambda: m.find_device(needles, q, out, 1 << 20, cnt, algo=algo), iters=20, warm=3, flush=flush)
                    print(f"fixed {n:9d} rows x {q:4d} {name:6s}: median {med:8.1f} us  best {best:8.1f} us")
            db.close()
    if "prof" in what:    # a few launches of the <=128-needle scan for ncu
        n = (10_000_000 // 2048 + 1) * 2048
        db = yp.CudaVectorDatabase(synthetic_rows=n)
        m = yp.CudaMatcher(db, 31)
        out = torch.zeros((1 << 20, 16), dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        for q in (8, 128):
            needles = torch.from_numpy(rng.integers(0, 256, size=(q, 32), dtype=np.uint8)).cuda()
            for algo in (yp.ALGO_TENSOR_PAIRED, yp.ALGO_TILED):
                for _ in range(2):
                    flush.sum()
                    m.find_device(needles, q, out, 1 << 20, cnt, algo=algo)
        torch.cuda.synchronize()
        db.close()
    if "hashpx" in what:   # hash kernel per pixel format
        k = yp.CudaKernel(0)
        n = 4096
        outs = {"hashes": torch.empty((n, 8, 32), dtype=torch.uint8, device="cuda"),
                "quality": torch.empty((n, 8), dtype=torch.float32, device="cuda"),
                "thresholds": torch.empty((n, 8), dtype=torch.float32, device="cuda")}
        for name, fmt, ch in (("LUMA8", yp.PIXEL_LUMA8, 1), ("RGB8", yp.PIXEL_RGB8, 3), ("RGBA8", yp.PIXEL_RGBA8, 4),
                              ("RGBA8+FMA", yp.PIXEL_RGBA8 | yp.PIXEL_FMA, 4)):
            frames = torch.empty((n, 512, 512) if ch == 1 else (n, 512, 512, ch), dtype=torch.uint8, device="cuda")
            yp.synth_pixels(frames, n, fmt)
            med, best = timeit(lambda: k.hash_batch(frames, out=outs, pixel_format=fmt), iters=10, flush=flush)
            print(f"hash {n} {name:9s} frames: median {med:9.1f} us  best {best:9.1f} us  {n / med:7.3f} Mimg/s  "
                  f"{n * 262144 * ch / med / 1e3:7.1f} GB/s")
            del frames
    if "tc4clk" in what:   # SM clock / power while the tensor matcher runs back to back for ~1.5 s
        import subprocess, tempfile
        n = (10_000_000 // 2048 + 1) * 2048
        db = yp.CudaVectorDatabase(synthetic_rows=n)
        m = yp.CudaMatcher(db, 31)
        out = torch.zeros((1 << 20, 16), dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        needles = torch.from_numpy(rng.integers(0, 256, size=(32768, 32), dtype=np.uint8)).cuda()
        for name, algo in (("mxf4", yp.ALGO_TENSOR), ("i8", yp.ALGO_TENSOR_I8), ("tiled", yp.ALGO_TILED)):
            f = tempfile.NamedTemporaryFile("w+", suffix=".csv")
            p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,"
                                  "clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                                  "--format=csv,noheader,nounits", "-lms", "50"], stdout=f)
            time.sleep(0.3)
            t0 = time.time()
            reps = 0
            while time.time() - t0 < 1.5:
                m.find_device(needles, 32768, out, 1 << 20, cnt, algo=algo)
                torch.cuda.synchronize()
                reps += 1
            dt = time.time() - t0
            p.terminate(); p.wait()
            f.seek(0)
            rowsv = [l.split(",") for l in f.read().splitlines() if l.strip()]
            clk = sorted(float(r[0]) for r in rowsv[4:])
            pw = sorted(float(r[1]) for r in rowsv[4:])
            cap = sum(1 for r in rowsv[4:] if "Active" in r[2] and "Not" not in r[2])
            print(f"{name}: {reps} launches in {dt:.2f} s = {dt / reps * 1e3:.2f} ms each, {n * 32768 * reps / dt / 1e12:.2f} Tpairs/s; "
                  f"SM clock median {clk[len(clk) // 2]:.0f} MHz (min {clk[0]:.0f}), power median {pw[len(pw) // 2]:.0f} W "
                  f"(max {pw[-1]:.0f}), sw_power_cap active in {cap}/{len(rowsv) - 4} samples")
        db.close()
    if "tiled" in what:
        n = (10_000_000 // 2048 + 1) * 2048
        db = yp.CudaVectorDatabase(synthetic_rows=n)
        m = yp.CudaMatcher(db, 31)
        out = torch.zeros((1 << 20, 16), dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        for q in (512, 1024, 4096, 32768):
            needles = torch.from_numpy(rng.integers(0, 256, size=(q, 32), dtype=np.uint8)).cuda()
            for name, algo in (("mxf4", yp.ALGO_TENSOR), ("i8", yp.ALGO_TENSOR_I8), ("tiled", yp.ALGO_TILED)):
                med, best = timeit(lambda: m.find_device(needles, q, out, 1 << 20, cnt, algo=algo), iters=3, warm=1)
                print(f"match 10M x {q:6d} {name:6s}: median {med / 1e3:9.2f} ms  {n * q / med / 1e6:8.3f} Tpairs/s  "
                      f"hits={int(cnt.item())}")
                cnt.zero_()
        db.close()
    if "hash" in what:
        k = yp.CudaKernel(0)
        for n in (592, 4096):
            frames = torch.empty((n, 512, 512), dtype=torch.float32, device="cuda")
            yp.synth_frames(frames, n, kind=1)
            outs = {"hashes": torch.empty((n, 8, 32), dtype=torch.uint8, device="cuda"),
                    "quality": torch.empty((n, 8), dtype=torch.float32, device="cuda"),
                    "thresholds": torch.empty((n, 8), dtype=torch.float32, device="cuda")}
            med, best = timeit(lambda: k.hash_batch(frames, out=outs), iters=10, flush=flush)
            print(f"hash {n:5d} frames: median {med:9.1f} us  best {best:9.1f} us  {n / med:7.3f} Mimg/s  "
                  f"{n * 1048576 / med / 1e3:7.1f} GB/s")
            del frames


if __name__ == "__main__":
    main()
