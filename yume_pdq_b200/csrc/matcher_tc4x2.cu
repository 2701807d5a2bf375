// matcher_tc4x2.cu -- tensor-core Hamming matcher, packed-FP4, TWO database rows per accumulator column.
//
// Same contract as matcher.cu (reference src/matching/mod.rs:336-362: every (row, needle) with
// popcount(row ^ needle) <= threshold, bit-exact).  The 128-bit prefix filter of matcher_tc4.cu leaves the tensor
// pipe half idle: its epilogue (one TMEM read + half a VIMNMX3 per pair) is the limiter.  Here one f32 accumulator
// carries the prefix dot products of one needle against TWO rows,
//     Z = 1.5 * 2^23  +  dot128(needle, row_lo)  +  65536 * dot128(needle, row_hi),
// built from five tcgen05.mma kind::mxf4 (K = 64) per tile: a constant "bias" MMA (1.5 x 2.0, A scale 2^22), two over
// row_hi's prefix with the needle tile scaled by 2^16 (ue8m0 block scale), two over row_lo's with scale 1.  All
// partial sums are integers below 2^24, so the f32 accumulation is exact (verified on a B200 by
// tools/tc4x2_test.cu), and the bit pattern of Z is
//     0x4B400000 + dot_lo + 65536 * dot_hi        (|dot_hi| < 64)
// i.e. two independent signed 16-bit fields: low half = dot_lo (two's complement), high half = 0x4B40 + dot_hi
// (- 1 when dot_lo < 0).  The epilogue tests both with ONE VIMNMX3.S16x2 per two accumulators (4 pairs), so a pair
// costs half a TMEM read and a quarter ALU instruction.  When |dot_hi| >= 64 the exponent of Z moves: for
// dot_hi >= 64 (high half 0x4B80 + (dot_hi - 64) / 2, low half dot_lo / 2) the column is always flagged, for
// dot_hi <= -64 the low half reads 2 * dot_lo, which is flagged whenever dot_lo would be.  Flagged columns are
// re-checked on the packed bits (both rows, full 256-bit distance), so the reported set is exact; no match can be lost.
//
// Persistent kernel, one CTA per SM, 12 warps: warps 0-7 epilogue (lane = needle, 120 columns each), warps 8-9 expand
// the CTA's next 960 database rows (first 16 bytes of each) into four 240-row e2m1 tiles (double buffered), warp 10
// streams the pre-expanded needle tiles (8 KB) through a 4-stage cp.async.bulk ring, warp 11 issues the MMAs.
// Shared memory: 2 x 60 KB (rows) + 4 x 8 KB (needles) + 11.5 KB (bias tiles).  TMEM: 2 x 240 accumulator columns +
// 4 x 8 columns of scale factors (one byte value per region, so the scale-factor layout is irrelevant).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc4_common.cuh"

namespace ypdq {
namespace {
using namespace tc4;

constexpr int X2_M = 128;                      // needles per tile (TMEM lanes)
constexpr int X2_N = 240;                      // accumulator columns per tile = rows per row tile
constexpr int X2_ACC = 2;                      // accumulator stages in TMEM
constexpr int X2_TILES = 2 * X2_ACC;           // row tiles per super-tile: accumulator t <- tiles 2t (low), 2t+1 (high)
constexpr int X2_ROWS = X2_TILES * X2_N;       // 960 rows per super-tile (needle tile reuse)
constexpr int X2_ROW_BYTES = 64;               // 128 e2m1 nibbles: the 128-bit prefix
constexpr int X2_A_BYTES = X2_M * X2_ROW_BYTES;      // 8 KB
constexpr int X2_B_BYTES = X2_N * X2_ROW_BYTES;      // 15 KB
constexpr int X2_SB_BYTES = X2_TILES * X2_B_BYTES;   // 60 KB per super-tile buffer
constexpr uint32_t X2_SBO = 8 * X2_ROW_BYTES;        // bytes between 8-row groups
constexpr uint32_t X2_LBO = 128;                     // bytes between K-adjacent 16-byte chunks
constexpr int X2_AC_BYTES = X2_M * 32, X2_BC_BYTES = X2_N * 32;   // bias tiles: K = 64 only
constexpr uint32_t X2_C_SBO = 8 * 32;
constexpr int X2_A_STAGES = 4;
constexpr int X2_EPI_WARPS = 8;
constexpr int X2_EXP_THREADS = 64;
constexpr int X2_WARP_PRODUCER = X2_EPI_WARPS + 2, X2_WARP_MMA = X2_EPI_WARPS + 3;
constexpr int X2_THREADS = 64 + X2_EXP_THREADS + 32 * X2_EPI_WARPS;
constexpr int X2_EB = X2_N / (X2_EPI_WARPS / 4);     // 120 accumulator columns per epilogue warp
static_assert(X2_EB == 120, "load_batch is written for 120 columns");
// TMEM columns of the scale-factor regions (8 columns each; a B region of 4 columns is too small for N = 240)
constexpr uint32_t X2_SF_B = 480, X2_SF_LO = 488, X2_SF_HI = 496, X2_SF_BIAS = 504;
constexpr uint32_t X2_BITS_BASE = 0x4B400000u;       // bit pattern of 1.5 * 2^23
constexpr size_t X2_SMEM = 2 * (size_t)X2_SB_BYTES + X2_A_STAGES * (size_t)X2_A_BYTES + X2_AC_BYTES + X2_BC_BYTES + 256;
constexpr uint32_t X2_IDESC = (1u << 7) | (1u << 10) | ((uint32_t)(X2_N >> 3) << 17) | (1u << 23) |
                              ((uint32_t)(X2_M >> 4) << 24);

__device__ __forceinline__ uint64_t x2_desc(uint32_t saddr, uint32_t sbo)
{
    // K-major, SWIZZLE_NONE canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units; version 1
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (uint64_t)(X2_LBO >> 4) << 16 | (uint64_t)(sbo >> 4) << 32 |
           (uint64_t)1 << 46;
}
__device__ __forceinline__ uint32_t x2_tile_off(uint32_t r, uint32_t kc)
{
    return (r >> 3) * X2_SBO + kc * X2_LBO + (r & 7u) * 16u;
}

// needles (q x 32 B) -> n_tiles tiles of 8 KB (first 128 bits, canonical layout); padding needles expand from zero bits
__global__ void __launch_bounds__(256)
expand_needles_x2_kernel(const uint4 *__restrict__ needles, uint32_t q, uint8_t *__restrict__ out, uint32_t n_tiles)
{
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n_tiles * X2_M)
        return;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (j < q)
        v = __ldg(needles + 2 * (size_t)j);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint8_t *tile = out + (size_t)(j / X2_M) * X2_A_BYTES;
    const uint32_t r = j % X2_M;
#pragma unroll
    for (int c = 0; c < 4; c++)
        *reinterpret_cast<uint4 *>(tile + x2_tile_off(r, c)) = expand32(w[c]);
}

__device__ __forceinline__ void x2_load_batch(int32_t (&v)[X2_EB], uint32_t taddr)
{
    T4_LD_X64(v, 0, taddr);
    T4_LD_X32(v, 64, taddr + 64);
    T4_LD_X16(v, 96, taddr + 96);
    T4_LD_X8(v, 112, taddr + 112);
}

// May some column of the batch hold a prefix dot product >= dot_min in either field?  lo_min = dot_min,
// hi_min = min(0x4B40 + dot_min - 1, 0x4B80) << 16.  No false negatives (see the header); 2 columns = 4 pairs per VIMNMX3.S16x2.
__device__ __forceinline__ bool x2_batch_any(const int32_t (&v)[X2_EB], int32_t lo_min, int32_t hi_min)
{
    uint32_t m4[4] = {(uint32_t)v[0], (uint32_t)v[1], (uint32_t)v[2], (uint32_t)v[3]};
#pragma unroll
    for (int j = 4; j < X2_EB; j += 2)
        m4[(j >> 1) & 3] = __vimax3_s16x2(m4[(j >> 1) & 3], (uint32_t)v[j], (uint32_t)v[j + 1]);
    const uint32_t m = __vmaxs2(__vimax3_s16x2(m4[0], m4[1], m4[2]), m4[3]);
    return ((int32_t)(int16_t)(m & 0xFFFFu) >= lo_min) | ((int32_t)m >= hi_min);
}

// rare path: verify both rows of every flagged column on the packed bits
__device__ __forceinline__ void x2_report(const int32_t (&v)[X2_EB], int32_t lo_min, int32_t hi_min, uint64_t row0,
                                          uint64_t n_rows, uint32_t needle, const uint4 *__restrict__ rows,
                                          const Tc4Sink &sink)
{
    uint32_t mask[(X2_EB + 31) / 32] = {};
#pragma unroll
    for (int j = 0; j < X2_EB; j++)
        mask[j >> 5] |= (uint32_t)(((int32_t)(int16_t)(v[j] & 0xFFFF) >= lo_min) | (v[j] >= hi_min)) << (j & 31);
    const uint4 b0 = __ldg(sink.needles_raw + 2 * (size_t)needle);
    const uint4 b1 = __ldg(sink.needles_raw + 2 * (size_t)needle + 1);
#pragma unroll
    for (int w = 0; w < (X2_EB + 31) / 32; w++) {
        uint32_t mw = mask[w];
#pragma unroll 1
        while (mw) {
            const uint64_t col_row = row0 + 32 * w + (__ffs(mw) - 1);
            mw &= mw - 1;
#pragma unroll 1
            for (uint32_t f = 0; f < 2; f++) {
                const uint64_t row = col_row + f * X2_N;     // low field: tile 2t, high field: tile 2t + 1
                if (row >= n_rows)
                    continue;
                const uint4 a0 = __ldg(rows + 2 * row), a1 = __ldg(rows + 2 * row + 1);
                const uint32_t dist = hamming256(a0, a1, b0, b1);
                if (dist <= sink.threshold)
                    tc4_emit(sink, row, needle, dist);
            }
        }
    }
}

__global__ void __launch_bounds__(X2_THREADS, 1)
match_tensor4x2_kernel(const uint4 *__restrict__ rows, uint64_t n_rows, const uint8_t *__restrict__ needles_exp,
                       uint32_t n_mtiles, uint32_t q, int32_t lo_min, int32_t hi_min, Tc4Sink sink)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sB = smem;                                      // 2 x 60 KB
    uint8_t *sA = smem + 2 * X2_SB_BYTES;                    // 4 x 8 KB
    uint8_t *sAc = sA + X2_A_STAGES * X2_A_BYTES;            // bias tiles
    uint8_t *sBc = sAc + X2_AC_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sBc + X2_BC_BYTES);
    uint64_t *a_full = bars, *a_empty = bars + 4, *b_full = bars + 8, *b_empty = bars + 10;
    uint64_t *d_full = bars + 12, *d_empty = bars + 16;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 20);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < X2_A_STAGES; i++) {
            bar_init(&a_full[i], 1);
            bar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            bar_init(&b_full[i], X2_EXP_THREADS);
            bar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < X2_ACC; i++) {
            bar_init(&d_full[i], 1);
            bar_init(&d_empty[i], X2_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // bias tiles: element 0 of every row = 1.5 (A, nibble 0x3) / 2.0 (B, nibble 0x4), the other 63 elements 0
    for (uint32_t i = threadIdx.x; i < (X2_AC_BYTES + X2_BC_BYTES) / 16; i += X2_THREADS) {
        const bool is_a = i < X2_AC_BYTES / 16;
        const uint32_t c = is_a ? i : i - X2_AC_BYTES / 16;   // 16-byte chunk: (c >> 4) = 8-row group, (c >> 3) & 1 = K chunk
        const uint32_t first = ((c >> 3) & 1u) == 0 ? (is_a ? 0x3u : 0x4u) : 0u;
        reinterpret_cast<uint4 *>(sAc)[i] = make_uint4(first, 0, 0, 0);
    }
    if (warp == X2_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // bias tiles: generic-proxy writes -> tensor-core reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (warp < 4) {
        // scale factors (ue8m0): rows and low-field needles 2^0, high-field needles 2^16, bias 2^22
        const uint32_t base = tmem + (((uint32_t)(warp & 3) * 32u) << 16);
#pragma unroll 1
        for (uint32_t c = X2_SF_B; c < 512; c++) {
            const uint32_t val = c < X2_SF_HI ? 0x7F7F7F7Fu : c < X2_SF_BIAS ? 0x8F8F8F8Fu : 0x95959595u;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(base + c), "r"(val) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const uint64_t n_stiles = (n_rows + X2_ROWS - 1) / X2_ROWS;

    if (warp == X2_WARP_PRODUCER) {
        // ===== needle tile producer =====
        if (lane == 0) {
            uint32_t it = 0;
            for (uint64_t st = blockIdx.x; st < n_stiles; st += gridDim.x)
                for (uint32_t m = 0; m < n_mtiles; m++, it++) {
                    const uint32_t s = it % X2_A_STAGES, use = it / X2_A_STAGES;
                    bar_wait(&a_empty[s], (use & 1u) ^ 1u);
                    bar_expect_tx(&a_full[s], X2_A_BYTES);
                    bulk_g2s(sA + (size_t)s * X2_A_BYTES, needles_exp + (size_t)m * X2_A_BYTES, X2_A_BYTES, &a_full[s]);
                }
        }
    } else if (warp == X2_WARP_MMA) {
        // ===== MMA issuer: 5 MMAs per (needle tile, accumulator) =====
        if (lane == 0) {
            const uint64_t d_bias_a = x2_desc(s32(sAc), X2_C_SBO), d_bias_b = x2_desc(s32(sBc), X2_C_SBO);
            const uint32_t sf_b = tmem + X2_SF_B;
            auto mma = [&](uint32_t d_addr, uint64_t da, uint64_t db, uint32_t sfa, uint32_t accumulate) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d_addr),
                    "l"(da), "l"(db), "r"(X2_IDESC), "r"(accumulate), "r"(sfa), "r"(sf_b)
                    : "memory");
            };
            uint32_t ia = 0, i = 0;
            for (uint64_t st = blockIdx.x; st < n_stiles; st += gridDim.x, i++) {
                const uint32_t b = i & 1u;
                bar_wait(&b_full[b], (i >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_addr = s32(sB + (size_t)b * X2_SB_BYTES);
                for (uint32_t m = 0; m < n_mtiles; m++, ia++) {
                    const uint32_t s = ia % X2_A_STAGES;
                    bar_wait(&a_full[s], (ia / X2_A_STAGES) & 1u);
                    const uint32_t a_addr = s32(sA + (size_t)s * X2_A_BYTES);
#pragma unroll
                    for (uint32_t t = 0; t < X2_ACC; t++) {
                        bar_wait(&d_empty[t], (ia & 1u) ^ 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t d_addr = tmem + t * X2_N;
#ifndef X2_EXPERIMENT_NO_BIAS     // developer timing experiment (wrong results): what would 4 MMAs per tile give?
                        mma(d_addr, d_bias_a, d_bias_b, tmem + X2_SF_BIAS, 0u);
#endif
#pragma unroll
                        for (uint32_t k = 0; k < 2; k++)
                            mma(d_addr, x2_desc(a_addr + k * 2 * X2_LBO, X2_SBO),
                                x2_desc(b_addr + (2 * t + 1) * X2_B_BYTES + k * 2 * X2_LBO, X2_SBO), tmem + X2_SF_HI,
#ifdef X2_EXPERIMENT_NO_BIAS
                                k);
#else
                                1u);
#endif
#pragma unroll
                        for (uint32_t k = 0; k < 2; k++)
                            mma(d_addr, x2_desc(a_addr + k * 2 * X2_LBO, X2_SBO),
                                x2_desc(b_addr + (2 * t) * X2_B_BYTES + k * 2 * X2_LBO, X2_SBO), tmem + X2_SF_LO, 1u);
                        umma_commit(&d_full[t]);      // accumulator ready for the epilogue
                    }
                    umma_commit(&a_empty[s]);         // needle stage may be refilled once these MMAs retire
                }
                umma_commit(&b_empty[b]);             // row buffer may be overwritten
            }
        }
    } else if (warp >= X2_EPI_WARPS) {
        // ===== row expanders: 64 threads, thread t expands rows t, t+64, ... of the super-tile, 4 rows at a time =====
        const uint32_t t = threadIdx.x - 32 * X2_EPI_WARPS;
        uint32_t i = 0;
        for (uint64_t st = blockIdx.x; st < n_stiles; st += gridDim.x, i++) {
            const uint32_t b = i & 1u;
            uint8_t *buf = sB + (size_t)b * X2_SB_BYTES;
#pragma unroll 1
            for (uint32_t g = 0; g < (X2_ROWS + 4 * X2_EXP_THREADS - 1) / (4 * X2_EXP_THREADS); g++) {
                uint4 lo[4];
#pragma unroll
                for (uint32_t h = 0; h < 4; h++) {  // loads first: the HBM latency overlaps the wait for the buffer
                    const uint32_t r = t + X2_EXP_THREADS * (4 * g + h);
                    const uint64_t row = st * X2_ROWS + r;
                    lo[h] = make_uint4(0, 0, 0, 0);
                    if (r < X2_ROWS && row < n_rows)
                        lo[h] = __ldg(rows + 2 * row);
                }
                if (g == 0)
                    bar_wait(&b_empty[b], ((i >> 1) & 1u) ^ 1u);
#pragma unroll
                for (uint32_t h = 0; h < 4; h++) {
                    const uint32_t r = t + X2_EXP_THREADS * (4 * g + h);
                    if (r < X2_ROWS) {
                        uint8_t *tile = buf + (r / X2_N) * X2_B_BYTES;
                        const uint32_t rr = r % X2_N;
                        const uint32_t w[4] = {lo[h].x, lo[h].y, lo[h].z, lo[h].w};
#pragma unroll
                        for (uint32_t kc = 0; kc < 4; kc++)
                            *reinterpret_cast<uint4 *>(tile + x2_tile_off(rr, kc)) = expand32(w[kc]);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
            bar_arrive(&b_full[b]);
        }
    } else {
        // ===== epilogue: TMEM lane quadrant = warp % 4, column half = warp / 4 =====
        const uint32_t quad = warp & 3u, part = (uint32_t)warp >> 2;
        const uint32_t lane_base = tmem + ((quad * 32u) << 16) + part * X2_EB;
        const uint32_t my_stiles =
            blockIdx.x < n_stiles ? (uint32_t)((n_stiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
        uint32_t dfull = s32(&d_full[0]), dempty = s32(&d_empty[0]), tbase = lane_base;
        int32_t lmin = lo_min, hmin = hi_min;
        asm volatile("" : "+r"(dfull), "+r"(dempty), "+r"(tbase), "+r"(lmin), "+r"(hmin));   // keep them in registers
        int32_t v[X2_EB];
        uint32_t parity = 0;
        for (uint32_t si = 0; si < my_stiles; si++) {
            for (uint32_t m = 0; m < n_mtiles; m++) {
#pragma unroll
                for (uint32_t acc = 0; acc < X2_ACC; acc++) {
                    bar_wait_addr(dfull + 8 * acc, parity);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    x2_load_batch(v, tbase + acc * X2_N);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    // the accumulator stage goes back to the MMA issuer before any arithmetic
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    if (lane == 0)
                        bar_arrive_addr(dempty + 8 * acc);
#ifdef X2_EXPERIMENT_NO_ARITH     // developer timing experiment: TMEM loads only
#pragma unroll
                    for (int j = 0; j < X2_EB; j++)
                        asm volatile("" ::"r"(v[j]));
                    if (false) {
#else
                    if (x2_batch_any(v, lmin, hmin)) {
#endif
                        // rare path: indices are rebuilt from opaque copies of the loop counters so that the
                        // compiler does not carry 64-bit row arithmetic through the hot loop
                        uint32_t si_ = si, m_ = m;
                        asm volatile("" : "+r"(si_), "+r"(m_));
                        const uint32_t needle = m_ * X2_M + quad * 32 + lane;
                        if (needle < q) {
                            const uint64_t st = blockIdx.x + (uint64_t)si_ * gridDim.x;
                            x2_report(v, lo_min, hi_min, st * X2_ROWS + acc * (2 * X2_N) + part * X2_EB, n_rows, needle,
                                      rows, sink);
                        }
                    }
                }
                parity ^= 1u;
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == X2_WARP_MMA)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace

// 128-bit prefix filter, two rows per accumulator column (threshold <= 63; candidates verified with POPC)
int launch_match_tensor4x2(const uint8_t *rows, size_t n_rows, uint64_t row_offset, const uint8_t *needles, size_t q,
                           uint32_t threshold, YumePDQ_Match *out, uint32_t cap, uint32_t *count,
                           unsigned long long *min_key, int mode, int sm_count, cudaStream_t s, uint32_t needle_base)
{
    if (n_rows == 0 || q == 0)
        return 0;
    if (threshold > 63)        // dot_min = 128 - 2 thr must stay positive
        return YUME_PDQ_CUDA_E_INVALID;
    Tc4Sink sink{out, cap, count, min_key, row_offset, (uint64_t)q, mode, needle_base,
                 reinterpret_cast<const uint4 *>(needles), threshold};
    const uint32_t n_mtiles = (uint32_t)((q + X2_M - 1) / X2_M);
    uint8_t *exp = nullptr;
    cudaError_t e = cudaMallocAsync(&exp, (size_t)n_mtiles * X2_A_BYTES, s);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? YUME_PDQ_CUDA_E_NOMEM : (int)e;
    }
    const uint32_t threads = n_mtiles * X2_M;
    expand_needles_x2_kernel<<<(threads + 255) / 256, 256, 0, s>>>(reinterpret_cast<const uint4 *>(needles), (uint32_t)q, exp,
                                                                   n_mtiles);
    YPDQ_LAUNCH_CHECK();
    const uint64_t n_stiles = (n_rows + X2_ROWS - 1) / X2_ROWS;
    const unsigned grid = (unsigned)std::min<uint64_t>(n_stiles, (uint64_t)sm_count);
    const int32_t dot_min = 128 - 2 * (int32_t)threshold;
    const int32_t lo_min = dot_min;
    // in the window the high half reads 0x4B40 + dot_hi (- 1 on a borrow); from dot_hi = 64 on the exponent of Z
    // moves and the half reads 0x4B80 + (dot_hi - 64) / 2 (and the low half dot_lo / 2): flag all of those
    const int32_t hi_min = (int32_t)(std::min<uint32_t>((X2_BITS_BASE >> 16) + (uint32_t)dot_min - 1u, 0x4B80u) << 16);
#ifdef X2_EXPERIMENT_NO_BIAS
    const int32_t lo_min_ = 40000, hi_min_ = 0x7FFF0000;     // nothing is ever flagged: timing only
#else
    const int32_t lo_min_ = lo_min, hi_min_ = hi_min;
#endif
    static bool attr_set = false;
    if (!attr_set) {
        YPDQ_CUDA_TRY(cudaFuncSetAttribute(match_tensor4x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)X2_SMEM));
        attr_set = true;
    }
    match_tensor4x2_kernel<<<grid, X2_THREADS, X2_SMEM, s>>>(reinterpret_cast<const uint4 *>(rows), n_rows, exp, n_mtiles,
                                                             (uint32_t)q, lo_min_, hi_min_, sink);
    YPDQ_LAUNCH_CHECK();
    YPDQ_CUDA_TRY(cudaFreeAsync(exp, s));
    return 0;
}

}  // namespace ypdq
