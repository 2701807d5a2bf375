// matcher_tc4x2.cu -- tensor-core Hamming matcher, packed-FP4, TWO database rows per accumulator column.
//
// Same contract as matcher.cu (reference src/matching/mod.rs:336-362: every (row, needle) with
// popcount(row ^ needle) <= threshold, bit-exact).  The 128-bit prefix filter of matcher_tc4.cu leaves the tensor
// pipe half idle: its epilogue (one TMEM read + half a VIMNMX3 per pair) is the limiter.  Here one f32 accumulator
// carries the prefix dot products of one needle against TWO rows,
//     Z = 1.5 * 2^23  +  dot_lo(needle, row_lo)  +  65536 * dot_hi(needle, row_hi),
// built from four tcgen05.mma kind::mxf4 (K = 64) per tile: two over row_hi's prefix with the needle tile scaled by
// 2^16 (ue8m0 block scale), two over row_lo's with scale 1.  Operands use the MASKED encoding of tc4_common.cuh
// (rows: hash bits ANDed into place as e2m1 {0, s}, five instructions per hash word; needles: +-1/s), so
//     dot(needle, row) = popc(needle) - hamming(needle, row)          over the bits a field covers,
// and a pair is a candidate when dot >= popc(needle) - threshold: a per-needle (= per TMEM lane) constant that
// expand_needles_x2_kernel precomputes.  The constant 1.5 * 2^23 = 192 * 2^16 rides in the high field's K range:
// its first six element positions (hash bits 0, 4, 8, 12, 16, 20 of word 0) hold constants instead of hash bits
// (rows: 6.0 each; needles: 6, 6, 6, 6, 6, 2 -> 5 * 36 + 12 = 192), so dot_lo covers 128 hash bits and dot_hi 122.
// All partial sums are integers below 2^24 (or multiples of 2^16), so the f32 accumulation is exact (verified on a
// B200 by tools/tc4x2_test.cu), and the bit pattern of Z is
//     0x4B400000 + dot_lo + 65536 * dot_hi        (|dot_hi| < 64)
// i.e. two independent signed 16-bit fields: low half = dot_lo (two's complement), high half = 0x4B40 + dot_hi
// (- 1 when dot_lo < 0).  The epilogue tests both with ONE VIMNMX3.S16x2 per two accumulators (4 pairs), so a pair
// costs half a TMEM read and a quarter ALU instruction.  When |dot_hi| >= 64 the exponent of Z moves: for
// dot_hi >= 64 (high half 0x4B80 + (dot_hi - 64) / 2, low half dot_lo / 2) the column is always flagged, for
// dot_hi <= -64 the low half reads 2 * dot_lo; the low-field minimum T is therefore lowered to 2 T when it is
// negative (needles with fewer set prefix bits than the threshold), so that both readings flag every dot_lo >= T.
// A prefix distance is a lower bound of the distance, and flagged columns are re-checked on the packed bits (both
// rows, full 256-bit distance), so the reported set is exact; no match can be lost.
//
// Persistent kernel, one CTA per SM, 16 warps.  A work unit is (960-row super-tile, chunk of needle tiles); units
// are dealt round-robin to the CTAs, chunk-major so that all CTAs stream the same needle tiles from L2.
//   warps 12-13 mask the unit's 960 rows (first 16 bytes of each) into 4 e2m1 row tiles (double buffered)
//   warp 14    streams the pre-expanded needle tiles (12 KB) through a 4-stage cp.async.bulk ring
//   warp 15    issues the MMAs: two accumulator stages of 240 TMEM columns, 4 MMAs (M 128, N 240, K 64) each
//   warps 0-11 epilogue: TMEM lane quadrant = warp % 4, column third = warp / 4: tcgen05.ld of 80 columns, hand the
//              stage back, then the max tree; flagged columns go to a candidate list that verify_candidates_kernel
//              checks on the packed bits after the scan.  LDTM moves 64 B/clk per sub-partition = 480 clk per 240-column stage,
//              the same as the MMA time of 4 x 120 clk; the two warps of a scheduler drift apart, so most of one
//              warp's max tree overlaps the other's TMEM read.  Tried and rejected on a B200: four stages of
//              128 / 112 columns with the warps of a scheduler on different stages (N = 128 MMAs re-read the needle
//              operand from shared memory twice as often and run out of its bandwidth: 16.9e12 pairs/s), and 64 + 56
//              column batches software-pipelined across tiles inside each warp (the stage is handed back later, which
//              stalls the MMAs: 24.1e12).
// Shared memory: 2 x 60 KB (rows) + 4 x 12 KB (needles).  TMEM: 480 accumulator columns + 3 x 8 columns of scale
// factors (one byte value per region, so the scale-factor layout is irrelevant).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc4_common.cuh"

#ifdef X2_TRACE
// developer build only (yume_pdq_b200/build.py variant trace X2_TRACE): per-role SM-clock timestamps of CTA 0
__device__ long long g_x2_trace[8][256];
extern "C" int yume_pdq_cuda_debug_x2_trace(long long *host_out)
{
    return (int)cudaMemcpyFromSymbol(host_out, g_x2_trace, sizeof(g_x2_trace));
}
#define X2_TR(role, slot)                                                       \
    do {                                                                        \
        if (blockIdx.x == 0 && (slot) < 256)                                    \
            g_x2_trace[role][slot] = clock64();                                 \
    } while (0)
#else
#define X2_TR(role, slot) do { } while (0)
#endif

namespace ypdq {
namespace {
using namespace tc4;

constexpr int X2_M = 128;                      // needles per tile (TMEM lanes)
constexpr int X2_STAGES = 2;                   // accumulator stages in TMEM
constexpr int X2_N = 240;                      // columns per stage
constexpr int X2_ROWS = 2 * X2_STAGES * X2_N;  // 960 rows per super-tile: every column carries two rows; stage t holds
                                               // rows [480 t, 480 t + 240) in its low and the next 240 in its high field
constexpr int X2_ROW_BYTES = 64;               // 128 e2m1 nibbles
constexpr int X2_SB_BYTES = X2_ROWS * X2_ROW_BYTES;   // 60 KB per super-tile buffer
constexpr uint32_t X2_SBO = 8 * X2_ROW_BYTES;  // row tiles: bytes between 8-row groups
constexpr uint32_t X2_LBO = 128;               // bytes between K-adjacent 16-byte chunks
// needle tile: per 8-needle group six 128-byte chunk blocks [c0', c1, c0, c1, c2, c3]; c0' = chunk 0 with the bias
// constants, used by the high field; every MMA reads two adjacent blocks
constexpr uint32_t X2_A_SBO = 6 * 128;
constexpr int X2_A_BYTES = (X2_M / 8) * X2_A_SBO;     // 12 KB
constexpr int X2_RAW_BYTES = X2_ROWS * 16;             // 15 KB: the 16-byte prefixes of a unit's rows as they lie in HBM
// Warp roles.  Three configurations of the same kernel:
//  * X2Large (12 epilogue warps x 80 columns, 2 row warps; 16 warps x 128 registers): many needle tiles per unit,
//    the epilogue is the bottleneck (8 / 12 / 16 epilogue warps: 29.2 / 29.9 / 28.2e12 pairs/s);
//  * X2Small (8 epilogue warps x 120 columns, 6 row warps): 2..7 needle tiles per unit, the row side has ~1000 clk
//    per tile to publish the next 960 rows.  The register file is re-split at run time (setmaxnreg): 176 per
//    epilogue thread, 80 for the rest;
//  * X2One: X2Small's roles for a single needle tile (see ONE_TILE).
// Warps [0, EPI) are epilogue warps (whole warpgroups); of the following warps two are the producer and the MMA
// issuer, the others mask rows.
// The packed rows reach shared memory by TMA bulk copies (producer lane -> RAW ring, 15 KB per unit) and are masked
// from there.  They must not be prefetched into registers: fence.proxy.async (needed between the masked st.shared
// and the MMAs) is MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC in SASS, and the membar waits for every global load the thread
// has in flight -- with register prefetching a <= 128-needle scan took MMA time PLUS HBM time (106 us for 10M rows;
// 83 us with the ring).  Also measured: two sets of row warps taking alternate units with plain loads (no ring; one
// set's round trip overlaps the other's masking): 85 us with two row buffers, 88 us with three -- the ~2500 clk
// round trip of a set, not the buffer chain, bounds it.  With the ring, shared-memory bandwidth does: per unit the
// MMAs fetch ~94 KB of operands, the masked stores are 61 KB, the ring 45 KB: ~200 KB at 128 B/clk = 1560 clk.
template <int EPI, int EXP, int ASTAGES, int RAWSTAGES, bool ONE = false>
struct X2Cfg {
    static constexpr int A_STAGES = ASTAGES;        // needle-tile ring slots (12 KB)
    static constexpr int RAW_STAGES = RAWSTAGES;    // row-prefix ring slots (15 KB)
    static constexpr int B_BUFS = 2;                // masked row buffers (60 KB): unit i lives in buffer i % B_BUFS
    static constexpr size_t SMEM = B_BUFS * (size_t)X2_SB_BYTES + ASTAGES * (size_t)X2_A_BYTES + RAWSTAGES * (size_t)X2_RAW_BYTES + 512;   // + 33 mbarriers and the TMEM address
    static_assert(SMEM <= 232448, "shared memory per CTA");
    // ONE_TILE: at most 128 needles.  The CTA expands the needle tile itself into shared memory (no expansion
    // kernel, no needle ring), every epilogue lane derives its needle's minima on its own, and the candidates are
    // verified by the warp that found them after its last unit (no verification kernel): ONE launch per scan.
    static constexpr bool ONE_TILE = ONE;
    static_assert(EPI % 4 == 0 && (EXP + 2) % 4 == 0, "whole warpgroups per role class");
    static constexpr int EPI_WARPS = EPI, EXP_WARPS = EXP, EXP_THREADS = 32 * EXP;
    static constexpr int WARP_PRODUCER = EPI + (EXP > 2 ? 4 : 2), WARP_MMA = WARP_PRODUCER + 1;
    static constexpr int THREADS = 32 * (EPI + EXP + 2);
    static constexpr int EB = 240 / (EPI / 4);            // accumulator columns per epilogue warp
    static constexpr bool SPLIT_REGS = EXP > 2;
    static constexpr int REGS_EPI = 176, REGS_OTHER = 80;
    static_assert(!SPLIT_REGS || 32 * (EPI * REGS_EPI + (EXP + 2) * REGS_OTHER) <= 65536 / THREADS / 8 * 8 * THREADS, "register pool");
};
using X2Large = X2Cfg<12, 2, 4, 2>;          // 120 + 48 + 30 KB
using X2Small = X2Cfg<8, 6, 4, 3>;           // 120 + 48 + 45 KB
using X2One = X2Cfg<8, 6, 1, 4, true>;       // 120 + 12 + 60 KB
// needle tiles (of 128) up to which X2Small is used; measured on 10M rows: 768 needles 0.36 ms (Small) vs 0.42 (Large),
// 1024: 0.44 vs 0.42, 1536: 0.61 vs 0.57, 2048: 0.80 vs 0.73
constexpr uint32_t X2_SMALL_MAX_TILES = 7;
constexpr uint32_t X2_SF_B = 480, X2_SF_LO = 488, X2_SF_HI = 496;   // TMEM columns of the scale-factor regions
constexpr uint32_t X2_BITS_BASE = 0x4B400000u;       // bit pattern of 1.5 * 2^23
constexpr int X2_BIAS_ELEMS = 6;                     // element positions of the high field that hold constants
constexpr uint32_t X2_ROW_BIAS = 0x00777777u, X2_NEEDLE_BIAS = 0x00477777u;   // e2m1 6.0 = 0x7, 2.0 = 0x4
// Units the producer lane prefetches into L2 ahead of the TMA row copies (cp.async.bulk.prefetch.L2, 15 KB each): the
// RAW ring alone keeps only 30..60 KB per SM in flight, which does not cover the HBM latency under load.
#ifndef X2_PREFETCH_UNITS_CFG
#define X2_PREFETCH_UNITS_CFG 6
#endif
constexpr uint32_t X2_PREFETCH_UNITS = X2_PREFETCH_UNITS_CFG;
#ifndef X2_MIN_CHUNK_TILES_CFG
#define X2_MIN_CHUNK_TILES_CFG 32
#define X2_UNITS_PER_SM_CFG 64
#endif
constexpr int X2_MIN_CHUNK_TILES = X2_MIN_CHUNK_TILES_CFG;   // needle tiles per work unit, at least (amortises the row expansion)
constexpr int X2_UNITS_PER_SM = X2_UNITS_PER_SM_CFG;         // work units per SM a small shard is cut into

// block-scaled instruction descriptor: A = B = e2m1, both K-major, ue8m0 scales, dense K = 64
constexpr uint32_t X2_IDESC = (1u << 7) | (1u << 10) | ((uint32_t)(X2_N >> 3) << 17) | (1u << 23) | ((uint32_t)(X2_M >> 4) << 24);

__device__ __forceinline__ uint64_t x2_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    // K-major, SWIZZLE_NONE canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units; version 1
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (uint64_t)(lbo >> 4) << 16 | (uint64_t)(sbo >> 4) << 32 | (uint64_t)1 << 46;
}

// needles (q x 32 B) -> n_tiles tiles of 12 KB in the needle-side MASKED encoding, plus the per-needle candidate
// minima {lo_min, hi_min} of the two accumulator fields (see the header; padding needles never flag)
constexpr uint32_t X2_BIAS_HASH_BITS = 0x00111111u;   // bits of hash word 0 the high field does not see

// candidate minima of the two accumulator fields for one needle (first 16 bytes in v)
__device__ __forceinline__ int2 x2_needle_mins(const uint4 &v, bool valid, int32_t threshold)
{
    if (!valid)
        return make_int2(32768, 0x7FFFFFFF);
    const int32_t p_lo = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    const int32_t t_lo = p_lo - threshold, t_hi = p_lo - __popc(v.x & X2_BIAS_HASH_BITS) - threshold;
    // in the window the high half reads 0x4B40 + dot_hi (- 1 on a borrow); from dot_hi = 64 on the exponent of Z
    // moves and the half reads 0x4B80 + (dot_hi - 64) / 2 (and the low half dot_lo / 2): flag all of those
    return make_int2(t_lo < 0 ? 2 * t_lo : t_lo, min((int32_t)(X2_BITS_BASE >> 16) + t_hi - 1, 0x4B80) << 16);
}

// needle r of a tile (first 16 bytes in v) -> its six 16-byte chunk slots [c0', c1, c0, c1, c2, c3]
__device__ __forceinline__ void x2_store_needle(uint8_t *tile, uint32_t r, const uint4 &v)
{
    uint8_t *grp = tile + (r >> 3) * X2_A_SBO + (r & 7u) * 16u;
    uint4 c0 = needle32(v.x);
    const uint4 c1 = needle32(v.y);
    *reinterpret_cast<uint4 *>(grp + 256) = c0;
    c0.x = (c0.x & 0xFF000000u) | X2_NEEDLE_BIAS;
    *reinterpret_cast<uint4 *>(grp) = c0;
    *reinterpret_cast<uint4 *>(grp + 128) = c1;
    *reinterpret_cast<uint4 *>(grp + 384) = c1;
    *reinterpret_cast<uint4 *>(grp + 512) = needle32(v.z);
    *reinterpret_cast<uint4 *>(grp + 640) = needle32(v.w);
}

__global__ void __launch_bounds__(256)
expand_needles_x2_kernel(const uint4 *__restrict__ needles, uint32_t q, uint8_t *__restrict__ out, int2 *__restrict__ mins,
                         uint32_t n_tiles, int32_t threshold)
{
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n_tiles * X2_M)
        return;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (j < q)
        v = __ldg(needles + 2 * (size_t)j);
    mins[j] = x2_needle_mins(v, j < q, threshold);
    x2_store_needle(out + (size_t)(j / X2_M) * X2_A_BYTES, j % X2_M, v);
}

template <int W>
__device__ __forceinline__ void x2_load(int32_t (&v)[W], uint32_t taddr)
{
    static_assert(W == 120 || W == 80 || W == 60, "120, 80 or 60 columns per warp");
    if constexpr (W == 80) {
        T4_LD_X64(v, 0, taddr);
        T4_LD_X16(v, 64, taddr + 64);
    } else if constexpr (W == 120) {
        T4_LD_X64(v, 0, taddr);
        T4_LD_X32(v, 64, taddr + 64);
        T4_LD_X16(v, 96, taddr + 96);
        T4_LD_X8(v, 112, taddr + 112);
    } else {
        T4_LD_X32(v, 0, taddr);
        T4_LD_X16(v, 32, taddr + 32);
        T4_LD_X8(v, 48, taddr + 48);
        T4_LD_X4(v, 56, taddr + 56);
    }
}

// May some column hold a prefix dot product that reaches this lane's (= needle's) minimum in either field?
// No false negatives (see the header); 2 columns = 4 pairs per VIMNMX3.S16x2.
template <int W>
__device__ __forceinline__ bool x2_any(const int32_t (&v)[W], int32_t lo_min, int32_t hi_min)
{
    uint32_t m4[4] = {(uint32_t)v[0], (uint32_t)v[1], (uint32_t)v[2], (uint32_t)v[3]};
#pragma unroll
    for (int j = 4; j < W; j += 2)
        m4[(j >> 1) & 3] = __vimax3_s16x2(m4[(j >> 1) & 3], (uint32_t)v[j], (uint32_t)v[j + 1]);
    const uint32_t m = __vmaxs2(__vimax3_s16x2(m4[0], m4[1], m4[2]), m4[3]);
    return ((int32_t)(int16_t)(m & 0xFFFFu) >= lo_min) | ((int32_t)m >= hi_min);
}

// Candidates of the prefix filter are not verified by the epilogue warp that finds them (an atomic or two dependent
// global loads per candidate would hold up the warp, and with it the accumulator stage, for a microsecond): every
// epilogue warp appends them to a private region of a list in global memory -- no atomics, the fill count lives in a
// register -- and verify_candidates_kernel works the regions off after the scan.  A warp whose region is full
// verifies inline, so nothing is ever dropped.
struct X2Candidate {
    uint64_t row;      // local row
    uint32_t needle;
    uint32_t pad;
};
#ifndef X2_CAND_SLOTS_CFG
#define X2_CAND_SLOTS_CFG 1024
#endif
constexpr uint32_t X2_CAND_SLOTS = X2_CAND_SLOTS_CFG;   // per epilogue warp
struct X2CandList {
    X2Candidate *items;     // [gridDim.x * X2_EPI_WARPS][X2_CAND_SLOTS]
    uint32_t *counts;       // [gridDim.x * X2_EPI_WARPS] candidates, then as many counts of flagged accumulator batches
    uint32_t n_regions;
};

__device__ __forceinline__ bool x2_verify(uint64_t row, uint32_t needle, const DbPlanes &rows, const Tc4Sink &sink)
{
    const uint4 b0 = __ldg(sink.needles_raw + 2 * (size_t)needle), b1 = __ldg(sink.needles_raw + 2 * (size_t)needle + 1);
    const uint4 a0 = __ldg(rows.pre + row), a1 = __ldg(rows.suf + row);
    const uint32_t dist = hamming256(a0, a1, b0, b1);
    if (dist <= sink.threshold)
        tc4_emit(sink, row, needle, dist);
    return dist <= sink.threshold;
}

// rare path, entered by the whole warp: both rows of every flagged column of every lane become candidates
template <int W>
__device__ __forceinline__ void x2_report(const int32_t (&v)[W], int32_t lo_min, int32_t hi_min, uint64_t row0, uint64_t n_rows,
                                          uint32_t needle, bool needle_ok, uint32_t lane, const DbPlanes &rows,
                                          const Tc4Sink &sink, X2Candidate *region, uint32_t &n_cand)
{
    static_assert(W % 8 == 0, "groups of 8 columns");
#pragma unroll
    for (int w = 0; w < W / 8; w++) {
        // group test first (4 instructions): a flagged warp spends most of this function skipping clean groups
        const uint32_t g = __vimax3_s16x2(__vimax3_s16x2((uint32_t)v[8 * w], (uint32_t)v[8 * w + 1], (uint32_t)v[8 * w + 2]),
                                          __vimax3_s16x2((uint32_t)v[8 * w + 3], (uint32_t)v[8 * w + 4], (uint32_t)v[8 * w + 5]),
                                          __vmaxs2((uint32_t)v[8 * w + 6], (uint32_t)v[8 * w + 7]));
        const bool flagged = needle_ok && (((int32_t)(int16_t)(g & 0xFFFFu) >= lo_min) | ((int32_t)g >= hi_min));
        if (!__any_sync(0xFFFFFFFFu, flagged))
            continue;
        uint32_t mw = 0;     // bit j: column 8 w + j of this lane is flagged
        if (flagged) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                mw |= (uint32_t)(((int32_t)(int16_t)(v[8 * w + j] & 0xFFFF) >= lo_min) | (v[8 * w + j] >= hi_min)) << j;
        }
#pragma unroll 1
        while (__any_sync(0xFFFFFFFFu, mw != 0)) {
            const bool have = mw != 0;
            const uint64_t col_row = row0 + 8 * w + (have ? __ffs(mw) - 1 : 0);
            mw &= mw - 1;
#pragma unroll 1
            for (uint32_t f = 0; f < 2; f++) {
                const uint64_t row = col_row + f * X2_N;  // low field: the stage's first 240 rows, high field: the next 240
                const bool cand = have && row < n_rows;
                const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, cand);
                const uint32_t slot = n_cand + __popc(ballot & ((1u << lane) - 1u));
                n_cand += __popc(ballot);
                if (cand) {
                    if (slot < X2_CAND_SLOTS)
                        region[slot] = X2Candidate{row, needle, 0u};
                    else
                        x2_verify(row, needle, rows, sink);
                }
            }
        }
    }
}

// one CTA per candidate region (= epilogue warp of the scan); also leaves the launch's count of flagged batches for the host
__global__ void __launch_bounds__(128)
verify_candidates_kernel(DbPlanes rows, X2CandList cand, Tc4Sink sink, MatchStats *stats, unsigned long long pairs,
                         unsigned long long *tally /* device: {flagged batches, matches, CTAs done}, zeroed */)
{
    const uint32_t n = min(cand.counts[blockIdx.x], X2_CAND_SLOTS);
    const X2Candidate *region = cand.items + (size_t)blockIdx.x * X2_CAND_SLOTS;
    uint32_t hits = 0;
    for (uint32_t i = threadIdx.x; i < n; i += 128)
        hits += x2_verify(region[i].row, region[i].needle, rows, sink);
    if (stats) {
        // statistics of the launch for ALGO_AUTO (matches verified inline after a list overflow are not counted)
        hits = __reduce_add_sync(0xFFFFFFFFu, hits);
        if ((threadIdx.x & 31) == 0 && hits)
            atomicAdd(&tally[1], (unsigned long long)hits);
        __syncthreads();
        if (threadIdx.x == 0) {
            atomicAdd(&tally[0], (unsigned long long)cand.counts[cand.n_regions + blockIdx.x]);
            __threadfence();
            if (atomicAdd(&tally[2], 1ull) == gridDim.x - 1) {     // last CTA: plain stores into the mapped host struct
                stats->flagged = atomicAdd(&tally[0], 0ull);
                stats->matches = atomicAdd(&tally[1], 0ull);
                __threadfence_system();
                stats->pairs = pairs;
            }
        }
    }
}

// Work units (super-tile, needle tile range), dealt round-robin to the CTAs in chunk-major order.  Every role of a
// CTA walks the same sequence with its own cursor (no division: the 64-bit divide of a direct decode cost 10 %).
struct X2Units {
    uint64_t n_stiles;
    uint32_t n_mtiles, chunk_tiles, n_chunks;
};
struct X2Cursor {
    uint64_t st;
    uint32_t ch;
    __device__ __forceinline__ explicit X2Cursor(const X2Units &u) : st(blockIdx.x), ch(0) { wrap(u); }
    __device__ __forceinline__ void wrap(const X2Units &u)
    {
        while (st >= u.n_stiles && ch < u.n_chunks) {
            st -= u.n_stiles;
            ch++;
        }
    }
    __device__ __forceinline__ bool valid(const X2Units &u) const { return ch < u.n_chunks; }
    __device__ __forceinline__ void next(const X2Units &u)
    {
        st += gridDim.x;
        wrap(u);
    }
    __device__ __forceinline__ uint32_t m0(const X2Units &u) const { return ch * u.chunk_tiles; }
    __device__ __forceinline__ uint32_t m1(const X2Units &u) const { return min(u.n_mtiles, (ch + 1) * u.chunk_tiles); }
};

// Epilogue of one warp: TMEM lanes 32 quad .. 32 quad + 31 (needles), columns EB part .. EB part + EB - 1 of both stages
template <class Cfg>
__device__ __forceinline__ void x2_epilogue(uint32_t tmem, uint32_t quad, uint32_t part, uint32_t lane, uint64_t *d_full,
                                            uint64_t *d_empty, const X2Units &units, uint32_t q, const int2 *__restrict__ mins,
                                            uint64_t n_rows, const DbPlanes &rows, const Tc4Sink &sink,
                                            const X2CandList &cand)
{
    constexpr int X2_EB = Cfg::EB;
    uint32_t dfull = s32(&d_full[0]), dempty = s32(&d_empty[0]);
    uint32_t tbase = tmem + ((quad * 32u) << 16) + part * X2_EB;
    asm volatile("" : "+r"(dfull), "+r"(dempty), "+r"(tbase));   // keep them in registers
    const int2 *my_mins = mins + quad * 32 + lane;               // this lane's needle within each tile
    uint32_t parity_idle = 0;
    int2 mn1 = make_int2(0, 0);
    if constexpr (Cfg::ONE_TILE) {
        if (quad * 32 >= q) {
            // no needle in this TMEM lane quadrant (q <= 32, 64, 96): only keep the accumulator hand-shake going,
            // so that the max trees of the empty lanes do not compete with the row warps for the ALU pipe
            for (X2Cursor c(units); c.valid(units); c.next(units)) {
#pragma unroll
                for (uint32_t acc = 0; acc < X2_STAGES; acc++) {
                    bar_wait_addr(dfull + 8 * acc, parity_idle);
                    if (lane == 0)
                        bar_arrive_addr(dempty + 8 * acc);
                }
                parity_idle ^= 1u;
            }
            return;
        }
        const uint32_t j = quad * 32 + lane;
        uint4 nv = make_uint4(0, 0, 0, 0);
        if (j < q)
            nv = __ldg(sink.needles_raw + 2 * (size_t)j);
        mn1 = x2_needle_mins(nv, j < q, (int32_t)sink.threshold);
    }
    const uint32_t region_id = blockIdx.x * Cfg::EPI_WARPS + part * 4 + quad;
    X2Candidate *region = cand.items + (size_t)region_id * X2_CAND_SLOTS;
    uint32_t n_cand = 0, n_flagged = 0;      // candidates listed / accumulator batches that took the rare path
    int32_t v[X2_EB];
    uint32_t parity = 0;
#ifdef X2_TRACE
    uint32_t trace_i = 0;
#endif
    for (X2Cursor c(units); c.valid(units); c.next(units)) {
        const uint64_t st = c.st;
        const uint32_t m1 = c.m1(units);
        for (uint32_t m = c.m0(units); m < m1; m++) {
            const int2 mn = Cfg::ONE_TILE ? mn1 : __ldg(my_mins + m * X2_M);   // issued before the wait: off the critical path
            const int32_t lmin = mn.x, hmin = mn.y;
#pragma unroll
            for (uint32_t acc = 0; acc < X2_STAGES; acc++) {
                bar_wait_addr(dfull + 8 * acc, parity);
#ifdef X2_TRACE
                const uint32_t trs = trace_i++;
                if (quad == 0 && part == 0 && lane == 0) X2_TR(4, trs);
#endif
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                x2_load<X2_EB>(v, tbase + acc * X2_N);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#ifdef X2_TRACE
                if (quad == 0 && part == 0 && lane == 0) X2_TR(5, trs);
#endif
                // the accumulator stage goes back to the MMA issuer before any arithmetic
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (lane == 0)
                    bar_arrive_addr(dempty + 8 * acc);
                if (__any_sync(0xFFFFFFFFu, x2_any<X2_EB>(v, lmin, hmin))) {
                    // rare path; opaque copies keep the 64-bit row arithmetic out of the hot loop
                    uint32_t m_ = m;
                    asm volatile("" : "+r"(m_));
                    n_flagged++;
                    const uint32_t needle = m_ * X2_M + quad * 32 + lane;
                    x2_report<X2_EB>(v, lmin, hmin, st * X2_ROWS + acc * (2 * X2_N) + part * X2_EB, n_rows, needle,
                                     needle < q, lane, rows, sink, region, n_cand);
                }
#ifdef X2_TRACE
                if (quad == 0 && part == 0 && lane == 0) X2_TR(6, trs);
#endif
            }
            parity ^= 1u;
        }
    }
    if constexpr (Cfg::ONE_TILE) {
        __syncwarp();     // the region was written by the other lanes of this warp
        const uint32_t n = min(n_cand, X2_CAND_SLOTS);
        for (uint32_t i = lane; i < n; i += 32)
            x2_verify(region[i].row, region[i].needle, rows, sink);
    } else if (lane == 0) {
        cand.counts[region_id] = n_cand;
        cand.counts[cand.n_regions + region_id] = n_flagged;
    }
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
match_tensor4x2_kernel(DbPlanes rows, uint64_t n_rows, const uint8_t *__restrict__ needles_exp,
                       X2Units units, uint32_t q, const int2 *__restrict__ mins, Tc4Sink sink, X2CandList cand)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int X2_A_STAGES = Cfg::A_STAGES, X2_RAW_STAGES = Cfg::RAW_STAGES;
    uint8_t *sB = smem;                                      // B_BUFS x 60 KB masked row tiles
    constexpr uint32_t NB = Cfg::B_BUFS;
    uint8_t *sA = smem + NB * X2_SB_BYTES;                   // A_STAGES x 12 KB needle tiles
    uint8_t *sR = sA + X2_A_STAGES * X2_A_BYTES;             // RAW_STAGES x 15 KB row prefixes
    uint64_t *bars = reinterpret_cast<uint64_t *>(sR + X2_RAW_STAGES * X2_RAW_BYTES);
    uint64_t *a_full = bars, *a_empty = bars + 4, *b_full = bars + 8, *b_empty = bars + 12;
    uint64_t *d_full = bars + 16, *d_empty = bars + 20, *r_full = bars + 24, *r_empty = bars + 28;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < X2_A_STAGES; i++) {
            bar_init(&a_full[i], 1);
            bar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < X2_RAW_STAGES; i++) {
            bar_init(&r_full[i], 1);
            bar_init(&r_empty[i], Cfg::EXP_THREADS);
        }
        for (int i = 0; i < (int)NB; i++) {
            bar_init(&b_full[i], Cfg::EXP_THREADS);
            bar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < X2_STAGES; i++) {
            bar_init(&d_full[i], 1);
            bar_init(&d_empty[i], Cfg::EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == Cfg::WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if constexpr (Cfg::ONE_TILE) {
        if (threadIdx.x < X2_M) {     // the one needle tile, resident in ring slot 0 for the whole kernel
            uint4 v = make_uint4(0, 0, 0, 0);
            if (threadIdx.x < q)
                v = __ldg(sink.needles_raw + 2 * (size_t)threadIdx.x);
            x2_store_needle(sA, threadIdx.x, v);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (warp < 4) {
        // scale factors (ue8m0): rows and low-field needles 2^0, high-field needles 2^16
        const uint32_t base = tmem + (((uint32_t)(warp & 3) * 32u) << 16);
#pragma unroll 1
        for (uint32_t c = X2_SF_B; c < 512; c++) {
            const uint32_t val = c < X2_SF_HI ? 0x7F7F7F7Fu : 0x8F8F8F8Fu;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(base + c), "r"(val) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (warp == Cfg::WARP_PRODUCER) {
        // ===== needle tile producer =====
        if constexpr (Cfg::SPLIT_REGS)
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_OTHER));
        auto unit_bytes = [&](uint64_t st) { return (uint32_t)min((uint64_t)X2_ROWS, n_rows - st * X2_ROWS) * 16u; };
        auto prefetch_unit = [&](uint64_t st) {
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(rows.pre + st * X2_ROWS), "r"(unit_bytes(st)) : "memory");
        };
        if (lane == 0) {
            // (1) L2 prefetch of the unit X2_PREFETCH_UNITS ahead, (2) TMA copy of the packed rows of the unit
            // RAW_STAGES ahead into the RAW ring (paced by the row warps' r_empty), (3) this unit's needle tiles
            X2Cursor ahead(units), raw(units);
            for (uint32_t d = 0; d < X2_PREFETCH_UNITS && ahead.valid(units); d++, ahead.next(units))
                prefetch_unit(ahead.st);
            uint32_t ir = 0;      // units handed to the RAW ring so far
            auto load_raw = [&]() {
                const uint32_t sl = ir % X2_RAW_STAGES, use = ir / X2_RAW_STAGES;
                bar_wait(&r_empty[sl], (use & 1u) ^ 1u);
                const uint32_t bytes = unit_bytes(raw.st);
                bar_expect_tx(&r_full[sl], bytes);
                bulk_g2s(sR + (size_t)sl * X2_RAW_BYTES, rows.pre + raw.st * X2_ROWS, bytes, &r_full[sl]);
                raw.next(units);
                ir++;
            };
            for (uint32_t d = 0; d < (uint32_t)X2_RAW_STAGES && raw.valid(units); d++)
                load_raw();
            uint32_t it = 0;
            for (X2Cursor c(units); c.valid(units); c.next(units)) {
                if (ahead.valid(units)) {
                    prefetch_unit(ahead.st);
                    ahead.next(units);
                }
                if (raw.valid(units))
                    load_raw();           // waits until the row warps have read unit (ir - RAW_STAGES) out of its slot
                if constexpr (Cfg::ONE_TILE)
                    continue;
                const uint32_t m1 = c.m1(units);
                for (uint32_t m = c.m0(units); m < m1; m++, it++) {
                    const uint32_t s = it % X2_A_STAGES, use = it / X2_A_STAGES;
                    bar_wait(&a_empty[s], (use & 1u) ^ 1u);
                    bar_expect_tx(&a_full[s], X2_A_BYTES);
                    bulk_g2s(sA + (size_t)s * X2_A_BYTES, needles_exp + (size_t)m * X2_A_BYTES, X2_A_BYTES, &a_full[s]);
                }
            }
        }
    } else if (warp == Cfg::WARP_MMA) {
        // ===== MMA issuer: 4 MMAs per (needle tile, accumulator stage) =====
        if constexpr (Cfg::SPLIT_REGS)
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_OTHER));
        if (lane == 0) {
            const uint32_t sf_b = tmem + X2_SF_B, sf_lo = tmem + X2_SF_LO, sf_hi = tmem + X2_SF_HI;
            auto mma = [&](uint32_t d_addr, uint64_t da, uint64_t db, uint32_t idesc, uint32_t sfa, uint32_t accumulate) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d_addr),
                    "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sf_b)
                    : "memory");
            };
            uint32_t ia = 0, i = 0;
            for (X2Cursor c(units); c.valid(units); c.next(units), i++) {
                const uint32_t m0 = c.m0(units), m1 = c.m1(units);
                const uint32_t b = i % NB;
                bar_wait(&b_full[b], (i / NB) & 1u);
                X2_TR(1, i);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_addr = s32(sB + (size_t)b * X2_SB_BYTES);
                for (uint32_t m = m0; m < m1; m++, ia++) {
                    const uint32_t s = Cfg::ONE_TILE ? 0u : ia % X2_A_STAGES;
                    if constexpr (!Cfg::ONE_TILE)
                        bar_wait(&a_full[s], (ia / X2_A_STAGES) & 1u);
                    const uint32_t a_addr = s32(sA + (size_t)s * X2_A_BYTES);
                    // needle operand: high field k = 0 reads [c0', c1], low field k = 0 [c0, c1], k = 1 [c2, c3]
                    const uint64_t a_hi0 = x2_desc(a_addr, 128, X2_A_SBO), a_lo0 = x2_desc(a_addr + 256, 128, X2_A_SBO);
                    const uint64_t a_k1 = x2_desc(a_addr + 512, 128, X2_A_SBO);
#pragma unroll
                    for (uint32_t t = 0; t < X2_STAGES; t++) {
                        bar_wait(&d_empty[t], (ia & 1u) ^ 1u);
                        X2_TR(3, 2 * ia + t);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t d_addr = tmem + t * X2_N, idesc = X2_IDESC;
                        // the stage's rows: 240 low rows, then 240 high rows
                        const uint32_t lo_tile = b_addr + t * (2 * X2_N * X2_ROW_BYTES);
                        const uint32_t hi_tile = lo_tile + X2_N * X2_ROW_BYTES;
                        mma(d_addr, a_hi0, x2_desc(hi_tile, X2_LBO, X2_SBO), idesc, sf_hi, 0u);
                        mma(d_addr, a_k1, x2_desc(hi_tile + 2 * X2_LBO, X2_LBO, X2_SBO), idesc, sf_hi, 1u);
                        mma(d_addr, a_lo0, x2_desc(lo_tile, X2_LBO, X2_SBO), idesc, sf_lo, 1u);
                        mma(d_addr, a_k1, x2_desc(lo_tile + 2 * X2_LBO, X2_LBO, X2_SBO), idesc, sf_lo, 1u);
                        umma_commit(&d_full[t]);      // accumulator ready for the epilogue
                        X2_TR(2, 2 * ia + t);
                    }
                    if constexpr (!Cfg::ONE_TILE)
                        umma_commit(&a_empty[s]);     // needle stage may be refilled once these MMAs retire
                }
                umma_commit(&b_empty[b]);             // row buffer may be overwritten
            }
        }
    } else if (warp >= Cfg::EPI_WARPS) {
        // ===== row warps: per unit, wait for its packed rows in the RAW ring and for a free row buffer, mask, publish
        if constexpr (Cfg::SPLIT_REGS)
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_OTHER));
        constexpr uint32_t X2_EXP_THREADS = Cfg::EXP_THREADS;
        static_assert(X2_ROWS % X2_EXP_THREADS == 0, "every row thread masks the same number of rows");
        constexpr uint32_t PER_THREAD = X2_ROWS / X2_EXP_THREADS;     // 15 or 5
        const uint32_t t = (uint32_t)(warp < Cfg::WARP_PRODUCER ? warp - Cfg::EPI_WARPS : warp - Cfg::EPI_WARPS - 2) * 32u + (uint32_t)lane;
        uint32_t i = 0;
        for (X2Cursor c(units); c.valid(units); c.next(units), i++) {
            const uint32_t b = i % NB, sl = i % X2_RAW_STAGES;
            uint8_t *buf = sB + (size_t)b * X2_SB_BYTES;
            const uint4 *src = reinterpret_cast<const uint4 *>(sR + (size_t)sl * X2_RAW_BYTES);
            const uint64_t left = n_rows - c.st * X2_ROWS;       // rows of this unit that exist (the others read as 0)
            bar_wait(&r_full[sl], (i / X2_RAW_STAGES) & 1u);
            uint4 rw[PER_THREAD];
#pragma unroll
            for (uint32_t h = 0; h < PER_THREAD; h++) {
                const uint32_t r = t + X2_EXP_THREADS * h;
                rw[h] = make_uint4(0, 0, 0, 0);
                if (r < left)
                    rw[h] = src[r];
            }
            bar_wait(&b_empty[b], ((i / NB) & 1u) ^ 1u);
            if (t == 0) X2_TR(0, i);
#pragma unroll
            for (uint32_t h = 0; h < PER_THREAD; h++) {
                const uint32_t r = t + X2_EXP_THREADS * h;
                // super-tile rows: stage 0 = [0, 480): 240 low + 240 high, stage 1 = [480, 960) the same; a
                // tile of 240 rows is 240 x 64 bytes in the canonical layout
                const uint32_t rr = r % X2_N;                                             // row within its tile
                const bool high = (r / X2_N) & 1u;
                uint8_t *tile = buf + (size_t)(r - rr) * X2_ROW_BYTES;                     // tiles are contiguous
                uint4 c0 = mask32(rw[h].x);
                if (high)
                    c0.x = (c0.x & 0xFF000000u) | X2_ROW_BIAS;
                uint8_t *p = tile + (rr >> 3) * X2_SBO + (rr & 7u) * 16u;
                *reinterpret_cast<uint4 *>(p) = c0;
                *reinterpret_cast<uint4 *>(p + X2_LBO) = mask32(rw[h].y);
                *reinterpret_cast<uint4 *>(p + 2 * X2_LBO) = mask32(rw[h].z);
                *reinterpret_cast<uint4 *>(p + 3 * X2_LBO) = mask32(rw[h].w);
            }
            bar_arrive(&r_empty[sl]);                                      // the slot's rows have been consumed
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
            bar_arrive(&b_full[b]);
            if (t == 0) X2_TR(7, i);
        }
    } else {
        if constexpr (Cfg::SPLIT_REGS)
            asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_EPI));
        x2_epilogue<Cfg>(tmem, (uint32_t)warp & 3u, (uint32_t)warp >> 2, (uint32_t)lane, d_full, d_empty, units, q, mins,
                    n_rows, rows, sink, cand);
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == Cfg::WARP_MMA)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace

// 128 / 122-bit prefix filter, two rows per accumulator column (threshold <= 60; candidates verified with POPC)
int launch_match_tensor4x2(DbPlanes rows, size_t n_rows, uint64_t row_offset, const uint8_t *needles, size_t q,
                           uint32_t threshold, YumePDQ_Match *out, uint32_t cap, uint32_t *count,
                           unsigned long long *min_key, int mode, int sm_count, cudaStream_t s, NeedleMap needle_base,
                           MatchStats *stats)
{
    if (n_rows == 0 || q == 0)
        return 0;
    if (threshold > 60)        // the high-field minimum 0x4B40 + popc - thr - 1 must stay inside the exponent window
        return YUME_PDQ_CUDA_E_INVALID;
    Tc4Sink sink{out, cap, count, min_key, row_offset, (uint64_t)q, mode, needle_base,
                 reinterpret_cast<const uint4 *>(needles), threshold};
    const uint32_t n_mtiles = (uint32_t)((q + X2_M - 1) / X2_M);
    // Work units: a small shard (a 10M-row DB split over 8 GPUs is 1302 super-tiles for 148 SMs) is cut along the
    // needle axis as well, until there are >= 64 units per SM or a unit is down to X2_MIN_CHUNK_TILES needle tiles.
    X2Units units;
    units.n_stiles = (n_rows + X2_ROWS - 1) / X2_ROWS;
    units.n_mtiles = n_mtiles;
    const uint64_t want_chunks = ((uint64_t)X2_UNITS_PER_SM * (uint64_t)sm_count + units.n_stiles - 1) / units.n_stiles;
    const uint64_t max_chunks = std::max<uint64_t>(1, n_mtiles / X2_MIN_CHUNK_TILES);
    const uint32_t n_chunks = (uint32_t)std::max<uint64_t>(1, std::min(want_chunks, max_chunks));
    units.chunk_tiles = (n_mtiles + n_chunks - 1) / n_chunks;
    units.n_chunks = (n_mtiles + units.chunk_tiles - 1) / units.chunk_tiles;
    const unsigned grid = (unsigned)std::min<uint64_t>(units.n_stiles * units.n_chunks, (uint64_t)sm_count);
    // one stream-ordered scratch block: expanded needle tiles + per-needle minima | candidate counts | candidate regions
    const bool one = n_mtiles == 1;      // <= 128 needles: everything in ONE launch (X2One)
    const size_t tile_bytes = one ? 0 : (size_t)n_mtiles * X2_A_BYTES;
    const size_t exp_bytes = tile_bytes + (one ? 0 : (size_t)n_mtiles * X2_M * sizeof(int2));
    const bool small = n_mtiles <= X2_SMALL_MAX_TILES;
    const size_t n_regions = (size_t)grid * (small ? X2Small::EPI_WARPS : X2Large::EPI_WARPS), count_bytes = (2 * n_regions * 4 + 255) / 256 * 256;
    uint8_t *exp = nullptr;
    const size_t scratch_bytes = exp_bytes + count_bytes + n_regions * X2_CAND_SLOTS * sizeof(X2Candidate);
    cudaError_t e = cudaMallocAsync(&exp, scratch_bytes + 32, s);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? YUME_PDQ_CUDA_E_NOMEM : (int)e;
    }
    X2CandList cand{reinterpret_cast<X2Candidate *>(exp + exp_bytes + count_bytes), reinterpret_cast<uint32_t *>(exp + exp_bytes), (uint32_t)n_regions};
    int2 *mins = reinterpret_cast<int2 *>(exp + tile_bytes);
    if (one) {
        static_assert(X2One::THREADS == X2Small::THREADS && X2One::EPI_WARPS == X2Small::EPI_WARPS, "same launch shape");
        YPDQ_CUDA_TRY(cudaFuncSetAttribute(match_tensor4x2_kernel<X2One>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)X2One::SMEM));
        match_tensor4x2_kernel<X2One><<<grid, X2One::THREADS, X2One::SMEM, s>>>(rows, n_rows, nullptr, units, (uint32_t)q, nullptr,
                                                                            sink, cand);
        YPDQ_LAUNCH_CHECK();
        YPDQ_CUDA_TRY(cudaFreeAsync(exp, s));
        return 0;
    }
    const uint32_t threads = n_mtiles * X2_M;
    expand_needles_x2_kernel<<<(threads + 255) / 256, 256, 0, s>>>(reinterpret_cast<const uint4 *>(needles), (uint32_t)q, exp,
                                                                   mins, n_mtiles, (int32_t)threshold);
    YPDQ_LAUNCH_CHECK();
    // per launch: function attributes are per device, and the call costs microseconds
    YPDQ_CUDA_TRY(cudaFuncSetAttribute(match_tensor4x2_kernel<X2Large>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)X2Large::SMEM));
    YPDQ_CUDA_TRY(cudaFuncSetAttribute(match_tensor4x2_kernel<X2Small>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)X2Small::SMEM));
    if (small)
        match_tensor4x2_kernel<X2Small><<<grid, X2Small::THREADS, X2Small::SMEM, s>>>(rows, n_rows, exp, units, (uint32_t)q, mins,
                                                                                sink, cand);
    else
        match_tensor4x2_kernel<X2Large><<<grid, X2Large::THREADS, X2Large::SMEM, s>>>(rows, n_rows, exp, units, (uint32_t)q, mins,
                                                                                sink, cand);
    YPDQ_LAUNCH_CHECK();
    unsigned long long *tally = reinterpret_cast<unsigned long long *>(exp + scratch_bytes);
    if (stats)
        YPDQ_CUDA_TRY(cudaMemsetAsync(tally, 0, 24, s));
    verify_candidates_kernel<<<(unsigned)n_regions, 128, 0, s>>>(rows, cand, sink, stats, (unsigned long long)n_rows * q, tally);
    YPDQ_LAUNCH_CHECK();
    YPDQ_CUDA_TRY(cudaFreeAsync(exp, s));
    return 0;
}

}  // namespace ypdq
