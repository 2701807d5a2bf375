// matcher.cu -- 100%-recall linear-scan Hamming matcher for sm_100a.
//
// Replaces the inner loops of CpuMatcher::scan/find (reference
// src/matching/mod.rs:232-257, 336-362; AVX-512 form 260-435) and of the WGSL
// shader of VulkanMatcher (src/matching/vulkan.rs:98-147): for every haystack
// row and every needle, dist = popcount(row ^ needle) over 256 bits; a match is
// dist <= threshold (inclusive).  All matches are reported, bit-exactly.
//
// B200 design notes:
//  * The obvious kernel spends 8 POPC per pair and POPC is a quarter-rate op
//    (16 lanes/clk/SM), which caps a 10M x 8 scan at ~140 us -- 3x the HBM time.
//    The kernels here first evaluate an EXACT lower bound on the distance from
//    the first 96 bits with a carry-save adder: s = x0^x1^x2, c = maj(x0,x1,x2),
//    popc(x0)+popc(x1)+popc(x2) = popc(s) + 2*popc(c)  -> 2 POPC + 4 LOP3 per
//    pair (x_i = row_i ^ needle_i; s comes from one LOP3 because row0^row1^row2
//    and needle0^needle1^needle2 are hoisted).  A pair can only match if that
//    prefix distance is <= threshold; survivors (3.4e-4 of random pairs at
//    threshold 31) get the full 8-word distance.  Speed is data dependent,
//    the match set is not.
//  * The database lies in HBM as TWO PLANES (DbPlanes, common.cuh): bytes 0..15 of
//    every row contiguously, then bytes 16..31.  The prefix filter reads the first
//    plane only -- 16 B per row, fully coalesced -- so a scan moves half the bytes
//    of the database; only the rare survivors touch the second plane.
//  * ALGO_TILED: rows in registers (4 per thread) x needles in shared memory
//    (2048 per CTA): the XU(POPC)-bound small-batch scan.
//  * ALGO_EXACT: unfiltered 8-POPC kernel (any threshold; the honest baseline).
//  * Hits are compacted with a warp ballot + one atomicAdd per warp.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace ypdq {

namespace {

// ------------------------------------------------------------- utilities --
__device__ __forceinline__ uint32_t lop3_carry(uint32_t x0, uint32_t x1, uint32_t s)
{
    // maj(x0, x1, x2) with x2 = s ^ x0 ^ x1, as one LOP3 over (x0, x1, s): LUT 0xD4
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xD4;" : "=r"(r) : "r"(x0), "r"(x1), "r"(s));
    return r;
}

// keeps ptxas from re-associating a hoisted xor back into the per-pair expression
__device__ __forceinline__ uint32_t opaque(uint32_t v)
{
    asm volatile("" : "+r"(v));
    return v;
}

__device__ __forceinline__ uint32_t prefix96_lower_bound(uint32_t r012, uint32_t r0, uint32_t r1,
                                                         uint32_t n012, uint32_t n0, uint32_t n1)
{
    uint32_t s = r012 ^ n012;
    uint32_t x0 = r0 ^ n0;
    uint32_t x1 = r1 ^ n1;
    uint32_t c = lop3_carry(x0, x1, s);
    return __popc(s) + 2u * __popc(c);
}

__device__ __forceinline__ uint32_t dist256(const uint4 &a0, const uint4 &a1, const uint4 &b0,
                                            const uint4 &b1)
{
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

struct Sink {
    YumePDQ_Match *out;
    uint32_t cap;
    uint32_t *count;
    unsigned long long *min_key;
    uint64_t row_offset;
    uint64_t q;
    int mode;
    NeedleMap needle_base;   // added to reported needle indices (LIST mode): pipelined callers match in chunks
};

// one hit from divergent code
__device__ __forceinline__ void emit_single(const Sink &k, uint64_t local_row, uint32_t needle,
                                            uint32_t dist)
{
    uint32_t pos = reserve_matches(k.count, 1u);
    if (k.mode == MODE_MIN) {
        atomicMin(k.min_key, (unsigned long long)(local_row * k.q + needle));
    } else if (pos < k.cap) {
        YumePDQ_Match m;
        m.row = local_row + k.row_offset;
        m.needle = k.needle_base(needle);
        m.dist = dist;
        k.out[pos] = m;
    }
}

// warp-ballot compaction: must be called by all 32 lanes of a warp
__device__ __forceinline__ void emit_ballot(const Sink &k, bool hit, uint64_t local_row,
                                            uint32_t needle, uint32_t dist)
{
    const unsigned full = 0xffffffffu;
    unsigned b = __ballot_sync(full, hit);
    if (b == 0)
        return;
    const unsigned lane = threadIdx.x & 31;
    const int leader = __ffs(b) - 1;
    uint32_t base = 0;
    if ((int)lane == leader)
        base = reserve_matches(k.count, (uint32_t)__popc(b));
    base = __shfl_sync(full, base, leader);
    if (hit) {
        if (k.mode == MODE_MIN) {
            atomicMin(k.min_key, (unsigned long long)(local_row * k.q + needle));
        } else {
            uint32_t pos = base + __popc(b & ((1u << lane) - 1u));
            if (pos < k.cap) {
                YumePDQ_Match m;
                m.row = local_row + k.row_offset;
                m.needle = k.needle_base(needle);
                m.dist = dist;
                k.out[pos] = m;
            }
        }
    }
}

// ------------------------------------------------------------ ALGO_EXACT --
constexpr int EX_THREADS = 256;
constexpr int EX_ROWS = 4;       // rows per thread
constexpr int EX_NCHUNK = 1024;  // needles per CTA (32 KB of shared memory)

__global__ void __launch_bounds__(EX_THREADS)
match_exact_kernel(DbPlanes rows, uint64_t n_rows, const uint4 *__restrict__ needles,
                   uint32_t q, uint32_t thr, Sink sink)
{
    __shared__ uint4 sn[EX_NCHUNK * 2];
    const uint32_t n_first = blockIdx.y * EX_NCHUNK;
    const uint32_t nc = min((uint32_t)EX_NCHUNK, q - n_first);
    for (uint32_t i = threadIdx.x; i < nc * 2; i += EX_THREADS)
        sn[i] = __ldg(needles + (size_t)n_first * 2 + i);
    const uint64_t tile = (uint64_t)blockIdx.x * (EX_THREADS * EX_ROWS);
    uint4 lo[EX_ROWS], hi[EX_ROWS];
    uint64_t ridx[EX_ROWS];
#pragma unroll
    for (int u = 0; u < EX_ROWS; u++) {
        ridx[u] = tile + (uint64_t)u * EX_THREADS + threadIdx.x;
        uint64_t r = ridx[u] < n_rows ? ridx[u] : 0;
        lo[u] = __ldg(rows.pre + r);
        hi[u] = __ldg(rows.suf + r);
    }
    __syncthreads();
    for (uint32_t i = 0; i < nc; i++) {
        const uint4 n0 = sn[2 * i], n1 = sn[2 * i + 1];
#pragma unroll
        for (int u = 0; u < EX_ROWS; u++) {
            uint32_t d = dist256(lo[u], hi[u], n0, n1);
            if (d <= thr && ridx[u] < n_rows)
                emit_single(sink, ridx[u], n_first + i, d);
        }
    }
}

// ------------------------------------------------------------ ALGO_TILED --
constexpr int TL_THREADS = 256;
constexpr int TL_NCHUNK = 2048;   // needles per CTA
constexpr int TL_SUB = 1024;      // needles between candidate-queue drains
constexpr int TL_QCAP = 2048;     // candidate queue entries

template <int TL_ROWS, int MINB>   // rows per thread held in registers, min resident CTAs per SM
__global__ void __launch_bounds__(TL_THREADS, MINB)
match_tiled_kernel(DbPlanes rows, uint64_t n_rows, const uint4 *__restrict__ needles,
                   uint32_t q, uint32_t thr, Sink sink)
{
    // dynamic shared memory sized by the launcher to the needle chunk actually present, so that small
    // batches leave room for many resident CTAs
    extern __shared__ uint4 tl_smem[];
    __shared__ uint32_t sqcount;
    const uint32_t n_alloc = min((uint32_t)TL_NCHUNK, q);
    uint4 *spre = tl_smem;                                             // {n0^n1^n2, n0, n1, n3}
    uint32_t *squeue = reinterpret_cast<uint32_t *>(tl_smem + n_alloc);  // (local row << 11) | local needle
    const uint32_t n_first = blockIdx.y * TL_NCHUNK;
    const uint32_t nc = min((uint32_t)TL_NCHUNK, q - n_first);
    for (uint32_t i = threadIdx.x; i < nc; i += TL_THREADS) {
        uint4 n = __ldg(needles + (size_t)(n_first + i) * 2);
        spre[i] = make_uint4(n.x ^ n.y ^ n.z, n.x, n.y, n.w);
    }
    if (threadIdx.x == 0)
        sqcount = 0;
    const uint64_t tile = (uint64_t)blockIdx.x * (TL_THREADS * TL_ROWS);
    uint32_t r012[TL_ROWS], r0[TL_ROWS], r1[TL_ROWS];
#pragma unroll
    for (int u = 0; u < TL_ROWS; u++) {
        uint64_t r = tile + (uint64_t)u * TL_THREADS + threadIdx.x;
        uint4 v = __ldg(rows.pre + (r < n_rows ? r : 0));
        r012[u] = opaque(v.x ^ v.y ^ v.z);
        r0[u] = v.x;
        r1[u] = v.y;
    }
    __syncthreads();

    for (uint32_t sub = 0; sub < nc; sub += TL_SUB) {
        const uint32_t sub_end = min(nc, sub + TL_SUB);
#pragma unroll 2
        for (uint32_t i = sub; i < sub_end; i++) {
            const uint4 np = spre[i];
            uint32_t lb[TL_ROWS];
            uint32_t m = 0xffffffffu;
#pragma unroll
            for (int u = 0; u < TL_ROWS; u++) {
                lb[u] = prefix96_lower_bound(r012[u], r0[u], r1[u], np.x, np.y, np.z);
                m = min(m, lb[u]);
            }
            if (m <= thr) {  // rare: some row of this thread survives the prefix filter
#pragma unroll
                for (int u = 0; u < TL_ROWS; u++) {
                    if (lb[u] <= thr) {
                        uint32_t lrow = (uint32_t)u * TL_THREADS + threadIdx.x;
                        uint32_t slot = atomicAdd(&sqcount, 1u);
                        if (slot < TL_QCAP) {
                            squeue[slot] = (lrow << 11) | i;
                        } else {  // queue full (adversarial data): verify in place
                            uint64_t r = tile + lrow;
                            if (r < n_rows) {
                                const uint4 a0 = __ldg(rows.pre + r), a1 = __ldg(rows.suf + r);
                                const uint4 *np2 = needles + (size_t)(n_first + i) * 2;
                                uint32_t d = dist256(a0, a1, __ldg(np2), __ldg(np2 + 1));
                                if (d <= thr)
                                    emit_single(sink, r, n_first + i, d);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        // drain the candidate queue: full 256-bit distance, warp-ballot compaction of hits
        const uint32_t cnt = min(sqcount, (uint32_t)TL_QCAP);
        for (uint32_t k = threadIdx.x; k < ((cnt + 31u) & ~31u); k += TL_THREADS) {
            bool hit = false;
            uint64_t r = 0;
            uint32_t nidx = 0, d = 0;
            if (k < cnt) {
                uint32_t code = squeue[k];
                r = tile + (code >> 11);
                nidx = n_first + (code & 2047u);
                if (r < n_rows) {
                    const uint4 a0 = __ldg(rows.pre + r), a1 = __ldg(rows.suf + r);
                    const uint4 *np2 = needles + (size_t)nidx * 2;
                    d = dist256(a0, a1, __ldg(np2), __ldg(np2 + 1));
                    hit = d <= thr;
                }
            }
            emit_ballot(sink, hit, r, nidx, d);
        }
        __syncthreads();
        if (threadIdx.x == 0)
            sqcount = 0;
        __syncthreads();
    }
}

}  // namespace

// --------------------------------------------------------------- launcher --
bool auto_picks_tensor_rows(size_t n_rows, size_t q, uint32_t threshold)
{
    return q <= 64 && threshold <= 40 && n_rows >= (q <= 8 ? 1000000u : 500000u);
}

int launch_match(DbPlanes rows, size_t n_rows, uint64_t row_offset, const uint8_t *needles,
                 size_t q, uint32_t threshold, YumePDQ_Match *out, uint32_t cap, uint32_t *count,
                 unsigned long long *min_key, int mode, int algo, int sm_count, cudaStream_t s,
                 NeedleMap needle_base, MatchStats *stats)
{
    if (n_rows == 0 || q == 0)
        return 0;
    if (q > 0xffffffffull)
        return YUME_PDQ_CUDA_E_INVALID;
    Sink sink{out, cap, count, min_key, row_offset, (uint64_t)q, mode, needle_base};
    const uint4 *needles4 = reinterpret_cast<const uint4 *>(needles);
    if (algo == YUME_PDQ_CUDA_ALGO_AUTO) {
        // Measured on B200 (10M rows, threshold 31; profiles/r02y, r03a quick.log): tiled POPC kernel 53 us for 8 needles,
        // 86 us for 16, 157 us for 32 (2.1e12 pairs/s from there on); rows-in-M tensor kernel 41 us up to 8 needles, 45 us
        // up to 16, 55 us up to 32, 70 us up to 64; needles-in-M (paired) tensor kernel 74..86 us for anything up to 128 needles (one
        // launch, X2One), 142 us for 256, 230 us for 512, 0.39 ms for 1024.  Above threshold 36 the large-batch tensor path
        // is the unpaired full kernel (1024 needles 0.8 ms vs tiled 4.7 ms); the 96-bit prefix bound of the POPC kernels
        // stops filtering once the threshold approaches 48.  A persistent tensor launch costs ~10 us before the first row:
        // small shards stay on the POPC kernel (568k rows x 8: 12.3 us either way, 142k rows: 9.2 vs 10.3).
        if (auto_picks_tensor_rows(n_rows, q, threshold))
            algo = YUME_PDQ_CUDA_ALGO_TENSOR_ROWS;
        else if (q >= (threshold <= 36 ? 14u : 512u) && (q >= 512 || n_rows >= 100000))
            algo = YUME_PDQ_CUDA_ALGO_TENSOR;
        else if (threshold > 40)
            algo = YUME_PDQ_CUDA_ALGO_EXACT;
        else
            algo = YUME_PDQ_CUDA_ALGO_TILED;
        // Data-dependent part (threshold <= 36, >= 512 needles): the paired prefix filter is the faster tensor kernel
        // unless many accumulator batches flag and few of their candidates are matches (MatchStats, common.cuh).  Every
        // paired launch leaves its counts in `stats` (mapped host memory).  Once a launch is past the break-even the
        // haystack is in FALLBACK: launches run the unfiltered 256-bit kernel, and every 8th .. 64th of them also runs
        // the filter on 1/64 of the rows (a slice that moves through the haystack; its matches are discarded) -- a slice
        // past the break-even doubles the interval, any other ends the fallback.
        if (algo == YUME_PDQ_CUDA_ALGO_TENSOR && threshold <= 36 && q >= 512 && stats) {
            const unsigned long long pairs_seen = stats->pairs;
            if (pairs_seen > 0) {
                const double flagged = (double)stats->flagged, matches = (double)stats->matches, pairs = (double)pairs_seen;
                const bool dense = flagged > kFlaggedBatchFraction * pairs / 5120.0 &&
                                   kFlaggedBatchNs * flagged - kUnfilteredMatchNs * matches > kUnfilteredExtraNsPerPair * pairs;
                stats->pairs = 0;
                if (dense) {
                    stats->probe_interval = stats->fallback ? std::min<uint32_t>(stats->probe_interval * 2, 64) : 8;
                    stats->fallback = 1;
                } else {
                    stats->fallback = 0;
                }
            }
            if (stats->fallback) {
                algo = YUME_PDQ_CUDA_ALGO_TENSOR_FULL;
                const size_t slice = std::min<size_t>(n_rows, std::max<size_t>(n_rows / 64, (size_t)960 * sm_count));
                if (++stats->since_probe >= stats->probe_interval && stats->probe_count && slice < n_rows) {
                    stats->since_probe = 0;
                    const size_t first = (size_t)((stats->probe_seq++ * 0x9E3779B97F4A7C15ull) % (n_rows - slice + 1));
                    int rc = launch_match_tensor4x2(DbPlanes{rows.pre + first, rows.suf + first}, slice, 0, needles, q, threshold,
                                                    nullptr, 0, stats->probe_count, nullptr, MODE_LIST, sm_count, s, 0, stats);
                    if (rc)
                        return rc;
                }
            }
        }
    }
    if (algo == YUME_PDQ_CUDA_ALGO_TENSOR_ROWS)
        return launch_match_tensor_rows(rows, n_rows, row_offset, needles, q, threshold, out, cap, count, min_key, mode,
                                        sm_count, s, needle_base);
    // candidates of the 128-bit prefix filter: 4.6e-9 of the random pairs at threshold 31, 6e-7 at 36, 1.7e-5 at 40
    if (algo == YUME_PDQ_CUDA_ALGO_TENSOR_PAIRED || (algo == YUME_PDQ_CUDA_ALGO_TENSOR && threshold <= 36))
        return launch_match_tensor4x2(rows, n_rows, row_offset, needles, q, threshold, out, cap, count, min_key,
                                      mode, sm_count, s, needle_base, stats);
    if (algo == YUME_PDQ_CUDA_ALGO_TENSOR && threshold <= 127)
        return launch_match_tensor4(rows, n_rows, row_offset, needles, q, threshold, out, cap, count, min_key,
                                    mode, sm_count, s, needle_base, -1);
    if (algo == YUME_PDQ_CUDA_ALGO_TENSOR_FULL || algo == YUME_PDQ_CUDA_ALGO_TENSOR_PREFIX)
        return launch_match_tensor4(rows, n_rows, row_offset, needles, q, threshold, out, cap, count, min_key,
                                    mode, sm_count, s, needle_base, algo == YUME_PDQ_CUDA_ALGO_TENSOR_PREFIX);
    if (algo == YUME_PDQ_CUDA_ALGO_TENSOR || algo == YUME_PDQ_CUDA_ALGO_TENSOR_I8)
        return launch_match_tensor(rows, n_rows, row_offset, needles, q, threshold, out, cap, count, min_key,
                                   mode, sm_count, s, needle_base);

    if (algo == YUME_PDQ_CUDA_ALGO_EXACT) {
        dim3 grid((unsigned)((n_rows + EX_THREADS * EX_ROWS - 1) / (EX_THREADS * EX_ROWS)),
                  (unsigned)((q + EX_NCHUNK - 1) / EX_NCHUNK));
        match_exact_kernel<<<grid, EX_THREADS, 0, s>>>(rows, n_rows, needles4, (uint32_t)q,
                                                       threshold, sink);
        YPDQ_LAUNCH_CHECK();
        return 0;
    }
    if (algo == YUME_PDQ_CUDA_ALGO_TILED) {
        // 4 rows per thread and >= 6 resident CTAs per SM overlap one CTA's HBM loads with the others' popcounts
        // (profiles/r01_scan_variants.log: 8 rows x 3 CTAs, 2 x 8 and 16 x 2 are all slower)
#ifndef TL_ROWS_CFG
#define TL_ROWS_CFG 4
#define TL_MINB_CFG 6
#endif
        constexpr int R = TL_ROWS_CFG, MINB = TL_MINB_CFG;
        dim3 grid((unsigned)((n_rows + TL_THREADS * R - 1) / (TL_THREADS * R)), (unsigned)((q + TL_NCHUNK - 1) / TL_NCHUNK));
        const size_t smem = std::min<size_t>(q, TL_NCHUNK) * sizeof(uint4) + TL_QCAP * sizeof(uint32_t);
        match_tiled_kernel<R, MINB><<<grid, TL_THREADS, smem, s>>>(rows, n_rows, needles4, (uint32_t)q, threshold, sink);
        YPDQ_LAUNCH_CHECK();
        return 0;
    }
    return YUME_PDQ_CUDA_E_INVALID;
}

}  // namespace ypdq
