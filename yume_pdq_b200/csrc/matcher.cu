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
//  * ALGO_STREAM: persistent CTAs, one per SM; a producer lane streams the AoS
//    database through an 8-stage ring of 16 KB shared-memory buffers with TMA
//    bulk copies (cp.async.bulk + mbarrier complete_tx); 16 consumer warps take
//    one row per lane per stage; the needle prefixes stay in registers.
//  * ALGO_TILED: rows in registers (8 per thread) x needles in shared memory
//    (2048 per CTA): the INT-pipe-bound regime for large needle batches; also
//    serves small batches with plain vectorised loads.
//  * ALGO_EXACT: unfiltered 8-POPC kernel (any threshold; the honest baseline).
//  * Hits are compacted with a warp ballot + one atomicAdd per warp.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

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
    uint32_t needle_base;   // added to reported needle indices (LIST mode): pipelined callers match in chunks
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
        m.needle = needle + k.needle_base;
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
                m.needle = needle + k.needle_base;
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
match_exact_kernel(const uint4 *__restrict__ rows, uint64_t n_rows, const uint4 *__restrict__ needles,
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
        lo[u] = __ldg(rows + 2 * r);
        hi[u] = __ldg(rows + 2 * r + 1);
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
constexpr int TL_DEFAULT_VARIANT = 3;  // 4 rows/thread, 6 CTAs/SM: fastest measured for q = 8 .. 256 (profiles/r01_scan_variants.log)

template <int TL_ROWS, int MINB>   // rows per thread held in registers, min resident CTAs per SM
__global__ void __launch_bounds__(TL_THREADS, MINB)
match_tiled_kernel(const uint4 *__restrict__ rows, uint64_t n_rows, const uint4 *__restrict__ needles,
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
        uint4 v = __ldg(rows + 2 * (r < n_rows ? r : 0));
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
                                const uint4 a0 = __ldg(rows + 2 * r), a1 = __ldg(rows + 2 * r + 1);
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
                    const uint4 a0 = __ldg(rows + 2 * r), a1 = __ldg(rows + 2 * r + 1);
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

// ----------------------------------------------------------- ALGO_STREAM --
constexpr int ST_STAGES = 8;
constexpr int ST_ROWS = 512;                   // rows per stage = consumer threads
constexpr int ST_STAGE_BYTES = ST_ROWS * 32;   // 16 KB
constexpr int ST_CONSUMER_WARPS = ST_ROWS / 32;
constexpr int ST_THREADS = ST_ROWS + 32;       // + one producer warp
constexpr int ST_MAXQ = 32;
constexpr size_t ST_SMEM = (size_t)ST_STAGES * ST_STAGE_BYTES + 2 * ST_STAGES * sizeof(uint64_t);

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                             uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

template <int NQ>  // needles rounded up to a multiple of 8 (8, 16, 24, 32)
__global__ void __launch_bounds__(ST_THREADS, 1)
match_stream_kernel(const uint8_t *__restrict__ rows, uint64_t n_rows,
                    const uint4 *__restrict__ needles, uint32_t q, uint32_t thr, Sink sink)
{
    // needle prefixes: registers for up to 16 needles, shared memory (broadcast LDS) beyond
    constexpr bool kRegs = NQ <= 16;
    __shared__ uint4 spre[kRegs ? 1 : NQ];
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + (size_t)ST_STAGES * ST_STAGE_BYTES);
    uint64_t *empty_bar = full_bar + ST_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < ST_STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], ST_CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint64_t n_chunks = (n_rows + ST_ROWS - 1) / ST_ROWS;

    if (warp == ST_CONSUMER_WARPS) {
        // ===== producer: one lane streams the database with TMA bulk copies =====
        if (lane == 0) {
            uint32_t it = 0;
            for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x, it++) {
                const uint32_t s = it % ST_STAGES, use = it / ST_STAGES;
                mbar_wait(&empty_bar[s], (use & 1u) ^ 1u);
                const uint64_t first = c * ST_ROWS;
                const uint32_t nr = (uint32_t)min((uint64_t)ST_ROWS, n_rows - first);
                mbar_arrive_expect_tx(&full_bar[s], nr * 32u);
                tma_bulk_g2s(smem + (size_t)s * ST_STAGE_BYTES, rows + first * 32u, nr * 32u,
                             &full_bar[s]);
            }
        }
        return;
    }

    // ===== consumers: one row per lane per stage =====
    uint32_t n012[kRegs ? NQ : 1], n0[kRegs ? NQ : 1], n1[kRegs ? NQ : 1];
    if constexpr (kRegs) {
#pragma unroll
        for (int j = 0; j < NQ; j++) {
            // padding slots repeat the last needle: their bound never undercuts a real one
            const uint4 n = __ldg(needles + 2 * min((uint32_t)j, q - 1));
            n012[j] = opaque(n.x ^ n.y ^ n.z);
            n0[j] = n.x;
            n1[j] = n.y;
        }
    } else {
        if (threadIdx.x < NQ) {
            const uint4 n = __ldg(needles + 2 * min((uint32_t)threadIdx.x, q - 1));
            spre[threadIdx.x] = make_uint4(n.x ^ n.y ^ n.z, n.x, n.y, n.w);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(ST_ROWS) : "memory");  // consumer warps only
    }
    uint32_t it = 0;
    for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x, it++) {
        const uint32_t s = it % ST_STAGES, use = it / ST_STAGES;
        mbar_wait(&full_bar[s], use & 1u);
        const uint64_t r = c * ST_ROWS + threadIdx.x;
        const uint4 *srow = reinterpret_cast<const uint4 *>(smem + (size_t)s * ST_STAGE_BYTES) +
                            2 * threadIdx.x;
        if (r < n_rows) {
            const uint4 lo = srow[0];
            const uint32_t r012 = opaque(lo.x ^ lo.y ^ lo.z);
            uint32_t m = 0xffffffffu;
#pragma unroll
            for (int j = 0; j < NQ; j++) {
                uint32_t lb;
                if constexpr (kRegs) {
                    lb = prefix96_lower_bound(r012, lo.x, lo.y, n012[j], n0[j], n1[j]);
                } else {
                    const uint4 pj = spre[j];
                    lb = prefix96_lower_bound(r012, lo.x, lo.y, pj.x, pj.y, pj.z);
                }
                m = min(m, lb);
            }
            if (m <= thr) {  // rare: verify every needle on the full row (still in shared memory)
                const uint4 hi = srow[1];
                for (uint32_t j = 0; j < q; j++) {
                    uint32_t d = dist256(lo, hi, __ldg(needles + 2 * j), __ldg(needles + 2 * j + 1));
                    if (d <= thr)
                        emit_single(sink, r, j, d);
                }
            }
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive(&empty_bar[s]);
    }
}

}  // namespace

// --------------------------------------------------------------- launcher --
int launch_match(const uint8_t *rows, size_t n_rows, uint64_t row_offset, const uint8_t *needles,
                 size_t q, uint32_t threshold, YumePDQ_Match *out, uint32_t cap, uint32_t *count,
                 unsigned long long *min_key, int mode, int algo, int sm_count, cudaStream_t s,
                 uint32_t needle_base)
{
    if (n_rows == 0 || q == 0)
        return 0;
    if (q > 0xffffffffull)
        return YUME_PDQ_CUDA_E_INVALID;
    Sink sink{out, cap, count, min_key, row_offset, (uint64_t)q, mode, needle_base};
    const uint4 *rows4 = reinterpret_cast<const uint4 *>(rows);
    const uint4 *needles4 = reinterpret_cast<const uint4 *>(needles);
    if (algo == YUME_PDQ_CUDA_ALGO_AUTO) {
        // Measured on B200 (10M rows, threshold 31; profiles/r02*/quick.log): tiled POPC kernel 54 us for 8 needles, 86 us
        // for 16, 157 us for 32 (2.1e12 pairs/s from there on); paired tensor kernel 74..86 us for anything up to 128
        // needles (one launch, X2One), 142 us for 256, 230 us for 512, 0.39 ms for 1024.  Above threshold 36 the tensor
        // path is the unpaired full kernel (1024 needles 0.8 ms vs tiled 4.7 ms); the 96-bit prefix bound of the POPC
        // kernels stops filtering once the threshold approaches 48.
        // (a persistent tensor launch costs ~12 us before the first row: not for tiny shards)
        if (q >= (threshold <= 36 ? 14u : 512u) && (q >= 512 || n_rows >= 100000))
            algo = YUME_PDQ_CUDA_ALGO_TENSOR;
        else if (threshold > 40)
            algo = YUME_PDQ_CUDA_ALGO_EXACT;
        else
            algo = YUME_PDQ_CUDA_ALGO_TILED;
    }
    if (algo == YUME_PDQ_CUDA_ALGO_STREAM && q > ST_MAXQ)
        algo = YUME_PDQ_CUDA_ALGO_TILED;
    // candidates of the 128-bit prefix filter: 4.6e-9 of the random pairs at threshold 31, 6e-7 at 36, 1.7e-5 at 40
    if (algo == YUME_PDQ_CUDA_ALGO_TENSOR_PAIRED || (algo == YUME_PDQ_CUDA_ALGO_TENSOR && threshold <= 36))
        return launch_match_tensor4x2(rows, n_rows, row_offset, needles, q, threshold, out, cap, count, min_key,
                                      mode, sm_count, s, needle_base);
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
        match_exact_kernel<<<grid, EX_THREADS, 0, s>>>(rows4, n_rows, needles4, (uint32_t)q,
                                                       threshold, sink);
        YPDQ_LAUNCH_CHECK();
        return 0;
    }
    if (algo == YUME_PDQ_CUDA_ALGO_TILED) {
        // 4 rows per thread and >= 6 resident CTAs per SM overlap one CTA's HBM loads with the others'
        // popcounts (10M rows x 8 needles: 55 us vs 74 us with 8 rows / 116 registers)
        static const int variant = [] {
            const char *e = getenv("YPDQ_TILED_VARIANT");   // developer knob for tools/quick_bench.py
            return e ? atoi(e) : -1;
        }();
        const int v = variant >= 0 ? variant : TL_DEFAULT_VARIANT;
#define LAUNCH_TILED(R, B)                                                                         \
    do {                                                                                          \
        dim3 grid((unsigned)((n_rows + TL_THREADS * R - 1) / (TL_THREADS * R)),                   \
                  (unsigned)((q + TL_NCHUNK - 1) / TL_NCHUNK));                                   \
        const size_t smem = std::min<size_t>(q, TL_NCHUNK) * sizeof(uint4) + TL_QCAP * sizeof(uint32_t); \
        match_tiled_kernel<R, B><<<grid, TL_THREADS, smem, s>>>(rows4, n_rows, needles4, (uint32_t)q, \
                                                             threshold, sink);                    \
    } while (0)
        switch (v) {
        case 1: LAUNCH_TILED(8, 4); break;
        case 2: LAUNCH_TILED(4, 4); break;
        case 3: LAUNCH_TILED(4, 6); break;
        case 4: LAUNCH_TILED(4, 8); break;
        case 5: LAUNCH_TILED(2, 8); break;
        case 6: LAUNCH_TILED(16, 2); break;
        case 7: LAUNCH_TILED(8, 3); break;
        default: LAUNCH_TILED(8, 1); break;
        }
#undef LAUNCH_TILED
        YPDQ_LAUNCH_CHECK();
        return 0;
    }
    if (algo == YUME_PDQ_CUDA_ALGO_STREAM) {
        const uint64_t n_chunks = (n_rows + ST_ROWS - 1) / ST_ROWS;
        unsigned grid = (unsigned)std::min<uint64_t>(n_chunks, (uint64_t)sm_count);
        const int nq8 = (int)((q + 7) / 8);
#define LAUNCH_STREAM(NQ)                                                                         \
    do {                                                                                          \
        YPDQ_CUDA_TRY(cudaFuncSetAttribute(match_stream_kernel<NQ>,                               \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                           (int)ST_SMEM));                                        \
        match_stream_kernel<NQ><<<grid, ST_THREADS, ST_SMEM, s>>>(rows, n_rows, needles4,         \
                                                                  (uint32_t)q, threshold, sink);  \
    } while (0)
        switch (nq8) {
        case 1: LAUNCH_STREAM(8); break;
        case 2: LAUNCH_STREAM(16); break;
        case 3: LAUNCH_STREAM(24); break;
        default: LAUNCH_STREAM(32); break;
        }
#undef LAUNCH_STREAM
        YPDQ_LAUNCH_CHECK();
        return 0;
    }
    return YUME_PDQ_CUDA_E_INVALID;
}

}  // namespace ypdq
