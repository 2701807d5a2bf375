// matcher_tc4.cu -- tensor-core Hamming matcher, packed-FP4 formulation (tcgen05.mma kind::mxf4).
//
// Same contract as matcher.cu (reference src/matching/mod.rs:336-362: every (row, needle) with
// popcount(row ^ needle) <= threshold, bit-exact).  Bits are mapped to e2m1 +-1.0 (nibbles 0x2 / 0xA), all
// ue8m0 block scales are 1.0, accumulation is f32 (exact: |dot| <= 256), so
//     dot(needle, row) = 256 - 2 * hamming(needle, row).
// Packed FP4 halves the operand bytes and doubles the MMA rate relative to the kind::i8 formulation of
// matcher_tc.cu (measured issue rate on B200: 8.9 POPS vs 4.55 POPS, tools/mma_peak.cu), i.e. a ceiling
// of 17.4e12 (row, needle) pairs/s instead of 8.9e12.
//
// Persistent kernel, one CTA per SM, 12 warps, warp-specialised:
//   warp 10      A producer: streams pre-expanded needle tiles (128 needles x 128 B = 16 KB, canonical
//                K-major SWIZZLE_NONE layout) from L2 into a 4-stage shared-memory ring with TMA bulk copies
//   warp 11      MMA issuer: per (needle tile, row tile) 4 x tcgen05.mma (M 128, N 240, K 64) into one of two
//                240-column TMEM accumulators; tcgen05.commit -> mbarriers
//   warps 8-9    B expanders: turn the CTA's next 480 database rows (15 KB of bits) into two 30 KB e2m1
//                row tiles (double buffered); each needle tile is streamed ONCE per 480 rows
//   warps 0-7    epilogue: tcgen05.ld (lane = needle, column = row; 120 columns per warp), 3-input max tree
//                over the raw f32 bits, rare hit path with exact distance = (256 - dot) / 2
// Shared memory: 2 x 60 KB (B) + 4 x 16 KB (A) = 184 KB.  TMEM: 2 x 240 accumulator columns + 32 columns
// of scale factors (every byte 0x7F = 2^0, so the scale-factor layout is irrelevant).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc4_common.cuh"

namespace ypdq {
namespace {
using namespace tc4;

#ifndef T4_N_CFG
#define T4_N_CFG 240
#endif
constexpr int T4_M = 128;                  // needles per tile (TMEM lanes)
constexpr int T4_N = T4_N_CFG;             // rows per tile (TMEM columns): 240 (2 accumulator stages) or 160 (3)
constexpr int T4_ROWS = 480;               // rows per super-tile (needle tile reuse)
constexpr int T4_NB = T4_ROWS / T4_N;      // row tiles per super-tile = accumulator stages in TMEM
static_assert(T4_N * T4_NB == T4_ROWS && T4_N % 16 == 0 && T4_NB >= 2 && T4_NB <= 4, "480 rows = NB tiles of N columns");
// KC = 16-byte K chunks (32 e2m1 elements each) per operand row: 8 = all 256 bits (exact dot product),
// 4 = the first 128 bits only (prefix filter: half the tensor work, candidates verified with POPC)
template <int KC>
struct T4Cfg {
    static constexpr int ROW_BYTES = KC * 16;              // 128 (256 e2m1 nibbles) or 64
    static constexpr int A_BYTES = T4_M * ROW_BYTES;       // 16 KB / 8 KB
    static constexpr int B_BYTES = T4_N * ROW_BYTES;       // 30 KB / 15 KB
    static constexpr int SB_BYTES = T4_NB * B_BYTES;       // 60 KB / 30 KB per super-tile buffer
    static constexpr uint32_t SBO = 8 * ROW_BYTES;         // bytes between 8-row groups
    static constexpr int KSTEPS = KC / 2;                  // tcgen05.mma K = 64 per instruction
    static constexpr int KBITS = KC * 32;
};
constexpr int T4_A_STAGES = 4;
constexpr uint32_t T4_PREFIX_MAX_THRESHOLD = 36;   // see launch_match_tensor4
#ifndef T4_EPI_WARPS_CFG
#define T4_EPI_WARPS_CFG 8
#endif
constexpr int T4_EPI_WARPS = T4_EPI_WARPS_CFG;   // 8: 2 per TMEM lane quadrant x 120 accumulator columns; 16: 4 x 60
constexpr int T4_EXP_THREADS = 64;          // B expander threads (the two warps after the epilogue warps)
// The warp scheduler favours the highest warp id among eligible warps, so the latency-critical single-thread
// roles get the highest ids: with the MMA issuer as warp 1 the epilogue warps sharing its scheduler delayed
// every tcgen05.mma issue by their own arithmetic (12.3e12 pairs/s).
constexpr int T4_WARP_PRODUCER = T4_EPI_WARPS + 2, T4_WARP_MMA = T4_EPI_WARPS + 3;
constexpr int T4_THREADS = 64 + T4_EXP_THREADS + 32 * T4_EPI_WARPS;  // 384 threads (up to 168 registers) or 640 (102)
constexpr uint32_t T4_LBO = 128;           // bytes between K-adjacent core matrices (16-byte chunks)
constexpr uint32_t T4_SF_COL = 480;        // TMEM columns 480..511: scale factors
template <int KC>
constexpr size_t t4_smem() { return 2 * (size_t)T4Cfg<KC>::SB_BYTES + T4_A_STAGES * (size_t)T4Cfg<KC>::A_BYTES + 256; }

template <int KC>
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr)
{
    // K-major, SWIZZLE_NONE canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units; version 1
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (uint64_t)(T4_LBO >> 4) << 16 |
           (uint64_t)(T4Cfg<KC>::SBO >> 4) << 32 | (uint64_t)1 << 46;
}
// block-scaled instruction descriptor: A = B = e2m1 (MXF4 format 1), both K-major, ue8m0 scales,
// N = 240, M = 128, K = 64 (dense), scale-factor ids 0
constexpr uint32_t T4_IDESC = (1u << 7) | (1u << 10) | ((uint32_t)(T4_N >> 3) << 17) | (1u << 23) |
                              ((uint32_t)(T4_M >> 4) << 24);

// canonical tile offset of (row r, 16-byte chunk kc)
template <int KC>
__device__ __forceinline__ uint32_t tile_off(uint32_t r, uint32_t kc)
{
    return (r >> 3) * T4Cfg<KC>::SBO + kc * T4_LBO + (r & 7u) * 16u;
}

// needles (q x 32 B) -> n_tiles tiles (16 KB, or 8 KB of the first 128 bits) in the canonical layout; padding
// needles expand from zero bits
template <int KC>
__global__ void __launch_bounds__(256)
expand_needles4_kernel(const uint4 *__restrict__ needles, uint32_t q, uint8_t *__restrict__ out, uint32_t n_tiles)
{
    constexpr uint32_t PARTS = KC / 4;                    // 16-byte parts of a needle that are expanded
    const uint32_t idx = blockIdx.x * 256 + threadIdx.x;  // one thread per (needle, part)
    const uint32_t j = idx / PARTS, half = idx % PARTS;
    if (j >= n_tiles * T4_M)
        return;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (j < q)
        v = __ldg(needles + 2 * (size_t)j + half);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint8_t *tile = out + (size_t)(j / T4_M) * T4Cfg<KC>::A_BYTES;
    const uint32_t r = j % T4_M;
#pragma unroll
    for (int c = 0; c < 4; c++)
        *reinterpret_cast<uint4 *>(tile + tile_off<KC>(r, 4 * half + c)) = expand32(w[c]);
}


constexpr int T4_EB = T4_N / (T4_EPI_WARPS / 4);   // accumulator columns per epilogue warp
static_assert(T4_EB == 60 || T4_EB == 120 || T4_EB == 80 || T4_EB == 40, "load_batch is written for 40/60/80/120 columns");

// T4_EB consecutive accumulator columns of this warp's 32 lanes -> v (asynchronous; needs wait::ld)
__device__ __forceinline__ void load_batch(int32_t (&v)[T4_EB], uint32_t taddr)
{
    if constexpr (T4_EB == 120) {
        T4_LD_X64(v, 0, taddr);
        T4_LD_X32(v, 64, taddr + 64);
        T4_LD_X16(v, 96, taddr + 96);
        T4_LD_X8(v, 112, taddr + 112);
    } else if constexpr (T4_EB == 80) {
        T4_LD_X64(v, 0, taddr);
        T4_LD_X16(v, 64, taddr + 64);
    } else if constexpr (T4_EB == 40) {
        T4_LD_X32(v, 0, taddr);
        T4_LD_X8(v, 32, taddr + 32);
    } else {
        T4_LD_X32(v, 0, taddr);
        T4_LD_X16(v, 32, taddr + 32);
        T4_LD_X8(v, 48, taddr + 48);
        T4_LD_X4(v, 56, taddr + 56);
    }
}

#ifndef T4_EPI_SQ_CFG
#define T4_EPI_SQ_CFG 0
#endif
// columns per batch whose test runs on the FMA pipe instead of the ALU pipe (0 = none): see batch_any
constexpr int T4_SQ = T4_EPI_SQ_CFG;
constexpr int T4_SQ_GROUP = 8;
static_assert(T4_SQ % T4_SQ_GROUP == 0 && T4_SQ < T4_EB, "whole groups");

// May some value of the batch reach min_bits (the f32 bit pattern of the smallest interesting dot product)?
// No false negatives; false positives only send the warp through report_hits, which re-checks every value.
//  * ALU pipe: 3-input integer max tree over the raw bits (positive floats order like integers), 2 values per
//    VIMNMX3;
//  * FMA pipe (the last T4_SQ columns, if any): sum of squares per group of 8, one FFMA per value; a value
//    >= dot_min makes its group's sum >= dot_min^2, and random groups almost never get there (chi-square tail:
//    4e-5 per group at 128 bits / threshold 31).  Both pipes issue at half rate, so splitting the batch between
//    them shortens the epilogue.
__device__ __forceinline__ bool batch_any(const int32_t (&v)[T4_EB], int32_t min_bits, float min_sq)
{
    constexpr int NMAX = T4_EB - T4_SQ;
    int32_t m4[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
    for (int j = 4; j < NMAX; j++)
        m4[j & 3] = max(m4[j & 3], v[j]);
    bool any = max(max(m4[0], m4[1]), max(m4[2], m4[3])) >= min_bits;
    if constexpr (T4_SQ > 0) {
        constexpr int NG = T4_SQ > 0 ? T4_SQ / T4_SQ_GROUP : 1;
        float sq[NG];
#pragma unroll
        for (int g = 0; g < T4_SQ / T4_SQ_GROUP; g++) {
            const float x = __int_as_float(v[NMAX + g * T4_SQ_GROUP]);
            sq[g] = x * x;
        }
#pragma unroll
        for (int e = 1; e < T4_SQ_GROUP; e++)
#pragma unroll
            for (int g = 0; g < T4_SQ / T4_SQ_GROUP; g++) {
                const float x = __int_as_float(v[NMAX + g * T4_SQ_GROUP + e]);
                sq[g] = fmaf(x, x, sq[g]);
            }
        float mx = sq[0];
#pragma unroll
        for (int g = 1; g < T4_SQ / T4_SQ_GROUP; g++)
            mx = fmaxf(mx, sq[g]);
        any |= mx >= min_sq;
    }
    return any;
}

// rare path: some row of this batch matches this needle (KC = 8), or agrees with it on enough of the first 128
// bits to be a candidate (KC = 4: the prefix distance is a lower bound of the distance, so no match is lost; the
// candidate's full 256-bit distance decides).  The unrolled part only collects a bit mask of the interesting
// columns; the loop over its set bits is not unrolled, so the rare path stays small.
template <int KC>
__device__ __forceinline__ void report_hits(const int32_t (&v)[T4_EB], int32_t min_bits, uint64_t row0, uint64_t n_rows,
                                            uint32_t needle, const DbPlanes &rows, const Tc4Sink &sink)
{
    if (KC == 8) {
#pragma unroll
        for (int j = 0; j < T4_EB; j++) {
            if (v[j] >= min_bits) {
                const uint64_t row = row0 + j;
                if (row < n_rows) {
                    const int32_t dot = (int32_t)__int_as_float(v[j]);
                    tc4_emit(sink, row, needle, (uint32_t)(256 - dot) >> 1);
                }
            }
        }
        return;
    }
    uint32_t mask[(T4_EB + 31) / 32] = {};
#pragma unroll
    for (int j = 0; j < T4_EB; j++)
        mask[j >> 5] |= (uint32_t)(v[j] >= min_bits) << (j & 31);
    const uint4 b0 = __ldg(sink.needles_raw + 2 * (size_t)needle);
    const uint4 b1 = __ldg(sink.needles_raw + 2 * (size_t)needle + 1);
#pragma unroll
    for (int w = 0; w < (T4_EB + 31) / 32; w++) {
        uint32_t mw = mask[w];
#pragma unroll 1
        while (mw) {
            const uint64_t row = row0 + 32 * w + (__ffs(mw) - 1);
            mw &= mw - 1;
            if (row >= n_rows)
                continue;
            const uint4 a0 = __ldg(rows.pre + row), a1 = __ldg(rows.suf + row);
            const uint32_t dist = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) +
                                  __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) +
                                  __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
            if (dist <= sink.threshold)
                tc4_emit(sink, row, needle, dist);
        }
    }
}

template <int KC>
__global__ void __launch_bounds__(T4_THREADS, 1)
match_tensor4_kernel(DbPlanes rows, uint64_t n_rows, const uint8_t *__restrict__ needles_exp,
                     uint32_t n_mtiles, uint32_t q, int32_t min_bits, Tc4Sink sink)
{
    using Cfg = T4Cfg<KC>;
    constexpr int T4_A_BYTES = Cfg::A_BYTES, T4_B_BYTES = Cfg::B_BYTES, T4_SB_BYTES = Cfg::SB_BYTES;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sB = smem;                                    // 2 x 60 KB (30 KB)
    uint8_t *sA = smem + 2 * T4_SB_BYTES;                  // 4 x 16 KB (8 KB)
    uint64_t *bars = reinterpret_cast<uint64_t *>(sA + T4_A_STAGES * T4_A_BYTES);
    uint64_t *a_full = bars, *a_empty = bars + 4, *b_full = bars + 8, *b_empty = bars + 10;
    uint64_t *d_full = bars + 12, *d_empty = bars + 16;          // T4_NB <= 4 accumulator stages
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 20);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < T4_A_STAGES; i++) {
            bar_init(&a_full[i], 1);
            bar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            bar_init(&b_full[i], T4_EXP_THREADS);
            bar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < T4_NB; i++) {
            bar_init(&d_full[i], 1);
            bar_init(&d_empty[i], T4_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == T4_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (warp < 4) {
        // scale factors: ue8m0 0x7F = 1.0 in every byte of columns 480..511 of all 128 lanes
        const uint32_t base = tmem + (((uint32_t)(warp & 3) * 32u) << 16) + T4_SF_COL;
        const uint32_t one = 0x7F7F7F7Fu;
#pragma unroll 1
        for (uint32_t c = 0; c < 32; c++)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(base + c), "r"(one) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const uint64_t n_stiles = (n_rows + T4_ROWS - 1) / T4_ROWS;

    if (warp == T4_WARP_PRODUCER) {
        // ===== A producer =====
        if (lane == 0) {
            uint32_t it = 0;
            for (uint64_t st = blockIdx.x; st < n_stiles; st += gridDim.x)
                for (uint32_t m = 0; m < n_mtiles; m++, it++) {
                    const uint32_t s = it % T4_A_STAGES, use = it / T4_A_STAGES;
                    bar_wait(&a_empty[s], (use & 1u) ^ 1u);
                    bar_expect_tx(&a_full[s], T4_A_BYTES);
                    bulk_g2s(sA + (size_t)s * T4_A_BYTES, needles_exp + (size_t)m * T4_A_BYTES, T4_A_BYTES, &a_full[s]);
                }
        }
    } else if (warp == T4_WARP_MMA) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t ia = 0, i = 0;
            for (uint64_t st = blockIdx.x; st < n_stiles; st += gridDim.x, i++) {
                const uint32_t b = i & 1u;
                bar_wait(&b_full[b], (i >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_addr = s32(sB + (size_t)b * T4_SB_BYTES);
                for (uint32_t m = 0; m < n_mtiles; m++, ia++) {
                    const uint32_t s = ia % T4_A_STAGES;
                    bar_wait(&a_full[s], (ia / T4_A_STAGES) & 1u);
                    const uint32_t a_addr = s32(sA + (size_t)s * T4_A_BYTES);
#pragma unroll
                    for (uint32_t t = 0; t < T4_NB; t++) {
                        const uint32_t acc = t;       // row tile t of every needle tile uses accumulator stage t
                        bar_wait(&d_empty[acc], (ia & 1u) ^ 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t d_addr = tmem + acc * T4_N;
#pragma unroll
                        for (uint32_t k = 0; k < (uint32_t)Cfg::KSTEPS; k++) {
                            const uint64_t da = smem_desc<KC>(a_addr + k * 2 * T4_LBO);
                            const uint64_t db = smem_desc<KC>(b_addr + t * T4_B_BYTES + k * 2 * T4_LBO);
                            const uint32_t accumulate = k > 0 ? 1u : 0u;
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t"
                                "setp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::
                                    "r"(d_addr),
                                "l"(da), "l"(db), "r"(T4_IDESC), "r"(accumulate), "r"(tmem + T4_SF_COL), "r"(tmem + T4_SF_COL + 16)
                                : "memory");
                        }
                        umma_commit(&d_full[acc]);    // accumulator ready for the epilogue
                    }
                    umma_commit(&a_empty[s]);         // A stage may be refilled once these MMAs retire
                }
                umma_commit(&b_empty[b]);             // B buffer may be overwritten
            }
        }
    } else if (warp >= T4_EPI_WARPS) {
        // ===== B expanders: 64 threads, thread t expands rows t, t+64, ... of the super-tile, 4 rows at a time =====
        const uint32_t t = threadIdx.x - 32 * T4_EPI_WARPS;
        uint32_t i = 0;
        for (uint64_t st = blockIdx.x; st < n_stiles; st += gridDim.x, i++) {
            const uint32_t b = i & 1u;
            uint8_t *buf = sB + (size_t)b * T4_SB_BYTES;
#pragma unroll 1
            for (uint32_t g = 0; g < 2; g++) {
                uint4 lo[4], hi[4];
#pragma unroll
                for (uint32_t h = 0; h < 4; h++) {  // loads first: the HBM latency overlaps the wait for the buffer
                    const uint32_t r = t + T4_EXP_THREADS * (4 * g + h);
                    const uint64_t row = st * T4_ROWS + r;
                    lo[h] = make_uint4(0, 0, 0, 0);
                    hi[h] = lo[h];
                    if (r < T4_ROWS && row < n_rows) {
                        lo[h] = __ldg(rows.pre + row);
                        if (KC == 8)
                            hi[h] = __ldg(rows.suf + row);
                    }
                }
                if (g == 0)
                    bar_wait(&b_empty[b], ((i >> 1) & 1u) ^ 1u);
#pragma unroll
                for (uint32_t h = 0; h < 4; h++) {
                    const uint32_t r = t + T4_EXP_THREADS * (4 * g + h);
                    if (r < T4_ROWS) {
                        uint8_t *tile = buf + (r / T4_N) * T4_B_BYTES;
                        const uint32_t rr = r % T4_N;
                        const uint32_t w[8] = {lo[h].x, lo[h].y, lo[h].z, lo[h].w, hi[h].x, hi[h].y, hi[h].z, hi[h].w};
#pragma unroll
                        for (uint32_t kc = 0; kc < (uint32_t)KC; kc++)
                            *reinterpret_cast<uint4 *>(tile + tile_off<KC>(rr, kc)) = expand32(w[kc]);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
            bar_arrive(&b_full[b]);
        }
    } else {
        // ===== epilogue: TMEM lane quadrant = warp % 4, column part = warp / 4 =====
        // 16 warps, 4 per scheduler: while one waits for its TMEM load the others run their max trees.
        const uint32_t quad = warp & 3u, part = (uint32_t)warp >> 2;
        const uint32_t lane_base = tmem + ((quad * 32u) << 16) + part * T4_EB;
        const uint32_t my_stiles =
            blockIdx.x < n_stiles ? (uint32_t)((n_stiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
        const uint32_t per_stile = n_mtiles * T4_NB;          // accumulator tiles per super-tile
        uint32_t dfull = s32(&d_full[0]), dempty = s32(&d_empty[0]), tbase = lane_base;
        int32_t mb = min_bits;
        float msq = __int_as_float(min_bits) * __int_as_float(min_bits);
        asm volatile("" : "+r"(dfull), "+r"(dempty), "+r"(tbase), "+r"(mb), "+f"(msq));   // keep them in registers (no rematerialisation)
        int32_t v[T4_EB];
        uint32_t parity = 0;
        for (uint32_t si = 0; si < my_stiles; si++) {
            for (uint32_t mt = 0; mt < per_stile; mt += T4_NB) {   // mt = m * T4_NB + t; accumulator stage = t
#pragma unroll
                for (uint32_t acc = 0; acc < T4_NB; acc++) {
                    bar_wait_addr(dfull + 8 * acc, parity);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    load_batch(v, tbase + acc * T4_N);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    // the accumulator stage goes back to the MMA issuer before any arithmetic
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    if (lane == 0)
                        bar_arrive_addr(dempty + 8 * acc);
                    if (batch_any(v, mb, msq)) {
                        // rare path: indices are rebuilt from opaque copies of the loop counters so that the
                        // compiler does not carry 64-bit row arithmetic through the hot loop
                        uint32_t si_ = si, mt_ = mt;
                        asm volatile("" : "+r"(si_), "+r"(mt_));
                        const uint32_t needle = (mt_ / T4_NB) * T4_M + quad * 32 + lane;
                        if (needle < q) {
                            const uint64_t st = blockIdx.x + (uint64_t)si_ * gridDim.x;
                            report_hits<KC>(v, min_bits, st * T4_ROWS + acc * T4_N + part * T4_EB, n_rows, needle, rows,
                                            sink);
                        }
                    }
                }
                parity ^= 1u;
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == T4_WARP_MMA)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int KC>
static int launch_tc4(DbPlanes rows, size_t n_rows, const uint8_t *needles, size_t q, uint32_t threshold,
                      const Tc4Sink &sink, int sm_count, cudaStream_t s)
{
    using Cfg = T4Cfg<KC>;
    const uint32_t n_mtiles = (uint32_t)((q + T4_M - 1) / T4_M);
    uint8_t *exp = nullptr;
    cudaError_t e = cudaMallocAsync(&exp, (size_t)n_mtiles * Cfg::A_BYTES, s);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? YUME_PDQ_CUDA_E_NOMEM : (int)e;
    }
    const uint32_t threads = n_mtiles * T4_M * (KC / 4);
    expand_needles4_kernel<KC><<<(threads + 255) / 256, 256, 0, s>>>(reinterpret_cast<const uint4 *>(needles), (uint32_t)q,
                                                                     exp, n_mtiles);
    YPDQ_LAUNCH_CHECK();
    const uint64_t n_stiles = (n_rows + T4_ROWS - 1) / T4_ROWS;
    const unsigned grid = (unsigned)std::min<uint64_t>(n_stiles, (uint64_t)sm_count);
    // a pair is inspected when its dot product over the KBITS bits reaches KBITS - 2 thr, i.e. when its distance
    // over those bits is <= thr
    const float dot_min = (float)(Cfg::KBITS - 2 * (int32_t)threshold);
    int32_t min_bits;
    memcpy(&min_bits, &dot_min, 4);
    // per launch: function attributes are per device, and the call costs microseconds
    YPDQ_CUDA_TRY(cudaFuncSetAttribute(match_tensor4_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)t4_smem<KC>()));
    match_tensor4_kernel<KC><<<grid, T4_THREADS, t4_smem<KC>(), s>>>(rows, n_rows, exp,
                                                                     n_mtiles, (uint32_t)q, min_bits, sink);
    YPDQ_LAUNCH_CHECK();
    YPDQ_CUDA_TRY(cudaFreeAsync(exp, s));
    return 0;
}

}  // namespace

// prefix = -1: automatic (prefix filter when the threshold makes candidates rare), 0: full 256-bit dot products,
// 1: 128-bit prefix filter
int launch_match_tensor4(DbPlanes rows, size_t n_rows, uint64_t row_offset, const uint8_t *needles, size_t q,
                         uint32_t threshold, YumePDQ_Match *out, uint32_t cap, uint32_t *count,
                         unsigned long long *min_key, int mode, int sm_count, cudaStream_t s, NeedleMap needle_base,
                         int prefix)
{
    if (n_rows == 0 || q == 0)
        return 0;
    if (threshold > 127)       // the raw-bit compare of the epilogue needs dot_min = 256 - 2 thr > 0
        return YUME_PDQ_CUDA_E_INVALID;
    // Prefix filter: P(Binomial(128, 1/2) <= thr) of the random pairs become candidates -- 4.6e-9 at thr = 31,
    // 6e-7 at 36, 1.7e-5 at 40, 0.45 at 63: beyond 36 the candidate checks would cost more than the tensor work saved.
    if (prefix < 0)
        prefix = threshold <= T4_PREFIX_MAX_THRESHOLD ? 1 : 0;
    if (prefix && threshold > 63)
        return YUME_PDQ_CUDA_E_INVALID;
    Tc4Sink sink{out, cap, count, min_key, row_offset, (uint64_t)q, mode, needle_base,
                 reinterpret_cast<const uint4 *>(needles), threshold};
    return prefix ? launch_tc4<4>(rows, n_rows, needles, q, threshold, sink, sm_count, s)
                  : launch_tc4<8>(rows, n_rows, needles, q, threshold, sink, sm_count, s);
}

}  // namespace ypdq
