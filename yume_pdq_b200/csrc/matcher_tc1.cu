// matcher_tc1.cu -- tensor-core Hamming scan for SMALL needle batches (1 .. 64 needles: the 8 dihedral hashes of one
// image, the reference's headline query): DATABASE ROWS IN THE M DIMENSION, operand A written to tensor memory
// straight from registers.
//
// Same contract as matcher.cu (reference src/matching/mod.rs:336-362: every (row, needle) with
// popcount(row ^ needle) <= threshold, bit-exact).  The POPC scan (match_tiled_kernel) is bound by the XU pipe (2 POPC
// per pair at 16 per clock and SM: 42 us for 10M rows x 8 needles + launch), the needles-in-M tensor scan
// (match_tensor4x2_kernel<X2One>) by shared-memory bandwidth (every row is written to and read back from shared memory
// as a 64-byte e2m1 operand: ~1.5 clk per row).  Here a row never touches shared memory as an operand:
//   * a worker thread owns one row of a 128-row tile = one TMEM lane.  It masks the row's 128-bit prefix into the
//     16 operand words of tc4_common.cuh (mask32: 20 integer instructions) and writes them to 16 TMEM columns with ONE
//     tcgen05.st.32x32b.x16 -- the A operand of tcgen05.mma kind::mxf4 may live in tensor memory;
//   * the needles are the B operand: N = 16 or 32 rows of 64 bytes in shared memory (needle32 encoding), built once per
//     launch; two MMAs (M 128, N, K 64) give D[row][needle] = threshold - hamming122(needle, row) as exact f32: six of
//     the 128 element positions (hash bits 0, 4 .. 20 of word 0) carry constants instead of hash bits -- rows 6, 6, 6, 6,
//     6, 1, needles the e2m1 decomposition of threshold - popc122(needle) (x1_bias_codes) -- so the candidate test is
//     the SIGN of the accumulator and needs no per-needle constant;
//   * the same thread reads its row's N accumulators back (one tcgen05.ld) and ANDs their bit patterns (3-input LOP3
//     tree): sign bit clear = some needle is within the threshold on the 122 prefix bits (6e-9 of the pairs on random
//     data at threshold 31); those pairs are verified on the packed 256 bits, so the reported set is exact.
// What bounds it (profiles/r02q/ncu_rows8_raw.csv, 10M rows x 8 needles in 41 us): the ALU pipe.  A row costs 22 integer
// instructions of operand masking + 4 of sign test + its share of the loop (two tcgen05.ld/st, an mbarrier wait and
// arrive, a cp.async): ~60 instructions, 80 % of them on the 64-lane ALU pipe, which ncu shows 58 % busy at an issue
// rate of 0.54 per scheduler.  HBM would need 25 us for the 16-byte prefix plane, the tensor pipe is 5 % busy.  One
// thread can issue a small-N MMA only every ~60 clk and a commit in ~54 (tools/x1_test.cu: a latency of the issuing
// thread -- four issuers reach four times the rate of one), hence one issuing warp per worker group.
//
// Persistent kernel, one CTA per SM: G groups of 4 worker warps (warp % 4 = TMEM lane quadrant) + one MMA-issuing warp each.
// Every worker thread streams ITS rows' 16-byte prefixes through a private ring of R cp.async copies (no producer warp,
// no barrier: a thread only reads what it copied itself; one 2 KB TMA bulk copy per tile was tried first and issues at
// ~430 clk per copy from one lane -- 137 us for 10M rows).  Every group owns S accumulator slots of 16 (A) + N (D)
// TMEM columns and keeps S - 1 tiles in flight.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "tc4_common.cuh"

#ifdef X1_TRACE
// developer build only (yume_pdq_b200/build.py variant trace1 X1_TRACE): per-role SM-clock timestamps of CTA 0, group 0
__device__ long long g_x1_trace[8][256];
extern "C" int yume_pdq_cuda_debug_x1_trace(long long *host_out)
{
    return (int)cudaMemcpyFromSymbol(host_out, g_x1_trace, sizeof(g_x1_trace));
}
#define X1_TR(role, slot)                                                       \
    do {                                                                        \
        if (blockIdx.x == 0 && (slot) < 256)                                    \
            g_x1_trace[role][slot] = clock64();                                 \
    } while (0)
#else
#define X1_TR(role, slot) do { } while (0)
#endif

namespace ypdq {
namespace {
using namespace tc4;

constexpr int X1_M = 128;                 // rows per tile (TMEM lanes)
constexpr uint32_t X1_SF = 496;           // TMEM columns 496..511: scale factors, all 1.0 (ue8m0 0x7F)

template <int N_, int T_, int G_, int S_, int R_>
struct X1Cfg {
    static constexpr int N = N_;          // needle slots (MMA N)
    static constexpr int T = T_;          // 128-row tiles per work unit: a worker thread owns T rows of a unit
    static constexpr int G = G_;          // worker groups of 128 threads (+ one MMA-issuing warp each)
    static constexpr int S = S_;          // accumulator slots per group (units in flight + the one being filled)
    static constexpr int R = R_;          // cp.async ring depth (units in flight per thread)
    static constexpr int SLOT_COLS = T * (16 + N);     // [A_0 .. A_{T-1} (16 columns each) | D_0 .. D_{T-1} (N each)]
    static_assert(G * S * SLOT_COLS <= (int)X1_SF, "TMEM columns");
    static constexpr int UNIT_ROWS = T * X1_M;
    static constexpr int WARP_MMA = 4 * G;       // first of G MMA-issuing warps
    static constexpr int THREADS = 32 * (4 * G + G);
    static constexpr int NB_BYTES = N * 64;
    static constexpr size_t RING_BYTES = (size_t)R * T * G * X1_M * 16;
    static constexpr size_t SMEM = NB_BYTES + 256 + RING_BYTES + 8 * (2 * G * S) + 16;
    static_assert(SMEM <= 232448, "shared memory per CTA");
    static constexpr uint32_t IDESC = (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((uint32_t)(X1_M >> 4) << 24);
};
#ifndef X1_G16_CFG
#define X1_T16_CFG 2
#define X1_G16_CFG 3
#define X1_S16_CFG 2
#define X1_T32_CFG 1
#define X1_G32_CFG 4
#define X1_S32_CFG 2
#define X1_R_CFG 8
#endif
#ifndef X1_ISSUER_SLEEP_NS
#define X1_ISSUER_SLEEP_NS 32
#endif
// Measured on a B200, 10M rows x {8, 16, 32} needles (profiles/r02y .. r03a; T tiles per unit, G groups, S slots):
//   N = 8:  T2 G4 S2 41.0 us, T3 G3 S2 41.1, T4 G2 S2 45.1        N = 16: T2 G3 S2 45.2, T3 G2 S2 49.2, T1 G4 S3 51.2
//   N = 32: T1 G4 S2 55.4, T1 G3 S3 59.5, T2 G2 S2 59.3           N = 64: T1 G3 S2 69.8
//   (needles-in-M kernel: 77 .. 80 us for any of them)
#ifndef X1_T8_CFG
#define X1_T8_CFG 2
#define X1_G8_CFG 4
#define X1_S8_CFG 2
#endif
using X1N8 = X1Cfg<8, X1_T8_CFG, X1_G8_CFG, X1_S8_CFG, X1_R_CFG>;
using X1N16 = X1Cfg<16, X1_T16_CFG, X1_G16_CFG, X1_S16_CFG, X1_R_CFG>;
using X1N32 = X1Cfg<32, X1_T32_CFG, X1_G32_CFG, X1_S32_CFG, X1_R_CFG>;
#ifndef X1_G64_CFG
#define X1_T64_CFG 1
#define X1_G64_CFG 3
#define X1_S64_CFG 2
#endif
using X1N64 = X1Cfg<64, X1_T64_CFG, X1_G64_CFG, X1_S64_CFG, X1_R_CFG>;

__device__ __forceinline__ uint64_t x1_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (uint64_t)(lbo >> 4) << 16 | (uint64_t)(sbo >> 4) << 32 | (uint64_t)1 << 46;
}

// units are dealt round-robin: iteration i of group g of CTA b works on unit (i * G + g) * gridDim.x + b
template <class Cfg>
__device__ __forceinline__ uint32_t x1_iterations(uint64_t n_units, uint32_t g)
{
    const uint64_t first = (uint64_t)g * gridDim.x + blockIdx.x, stride = (uint64_t)Cfg::G * gridDim.x;
    return first < n_units ? (uint32_t)((n_units - first + stride - 1) / stride) : 0u;
}

// mbarrier wait WITHOUT a suspend-time hint: with the 2 us hint of tc4::bar_wait a parked waiter woke ~4400 clk after the
// phase had completed (traced on a B200), which is the whole latency budget of this pipeline
__device__ __forceinline__ void x1_wait(uint32_t addr, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "X1_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra X1_DONE;\n\t"
        "bra X1_WAIT;\n"
        "X1_DONE:\n\t}" ::"r"(addr),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ bool bar_try_addr_nohint(uint32_t addr, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    return done != 0;
}

// rare path: a pair that passed the prefix filter is verified on the packed 256 bits
__device__ __noinline__ void x1_verify(uint64_t row, uint32_t needle, DbPlanes rows, const Tc4Sink &sink)
{
    const uint4 a0 = __ldg(rows.pre + row), a1 = __ldg(rows.suf + row);
    // (the needles of a mailbox query were written during this launch: L2 loads, not the read-only path)
    const uint4 b0 = __ldcg(sink.needles_raw + 2 * (size_t)needle), b1 = __ldcg(sink.needles_raw + 2 * (size_t)needle + 1);
    const uint32_t dist = hamming256(a0, a1, b0, b1);
    if (dist <= sink.threshold)
        tc4_emit(sink, row, needle, dist);
}

constexpr uint32_t X1_ROW_BIAS = 0x00277777u;      // row side of the six constant positions: 6, 6, 6, 6, 6 and 1 (e2m1 0x7, 0x2)
constexpr uint32_t X1_BIAS_HASH_BITS = 0x00111111u;  // bits of hash word 0 the filter does not see

// e2m1 codes e_0 .. e_5 (one per nibble) with 6 * (e_0 + .. + e_4) + 1 * e_5 = bias, for any integer |bias| <= 180
__device__ __forceinline__ uint32_t x1_bias_codes(int32_t bias)
{
    int32_t e5 = bias % 3;                    // -2 .. 2  ->  -1, 0, 1 (bias - e5 divisible by 3)
    e5 = e5 == 2 ? -1 : (e5 == -2 ? 1 : e5);
    int32_t h = (bias - e5) / 3;              // = 2 * (e_0 + .. + e_4): an integer number of halves
    const uint32_t sign = h < 0 ? 8u : 0u;
    h = h < 0 ? -h : h;
    uint32_t codes = e5 == 0 ? 0u : ((e5 < 0 ? 8u : 0u) | 2u);      // +-1.0
    codes <<= 20;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        // largest e2m1 magnitude (in halves: 12, 8, 6, 4, 3, 2, 1 = codes 7 .. 1) not above what is left
        uint32_t c = 0;
        int32_t take = 0;
        if (h >= 12) { c = 7; take = 12; }
        else if (h >= 8) { c = 6; take = 8; }
        else if (h >= 6) { c = 5; take = 6; }
        else if (h >= 4) { c = 4; take = 4; }
        else if (h >= 3) { c = 3; take = 3; }
        else if (h >= 2) { c = 2; take = 2; }
        else if (h >= 1) { c = 1; take = 1; }
        h -= take;
        codes |= (c ? (c | sign) : 0u) << (4 * k);
    }
    return codes;
}

template <int N>
__device__ __forceinline__ uint32_t x1_and_tree(const uint32_t (&d)[N])
{
    if constexpr (N == 1) {
        return d[0];
    } else {
        constexpr int M = (N + 2) / 3;
        uint32_t e[M];
#pragma unroll
        for (int o = 0; o < M; o++)
            e[o] = d[3 * o] & (3 * o + 1 < N ? d[3 * o + 1 < N ? 3 * o + 1 : 0] : 0xFFFFFFFFu) &
                   (3 * o + 2 < N ? d[3 * o + 2 < N ? 3 * o + 2 : 0] : 0xFFFFFFFFu);
        return x1_and_tree<M>(e);
    }
}

struct NoHook {};

// Hook = MailboxHook: the sharded query's whole exchange rides in this launch (api.cu, "peer mailboxes").  The root's
// CTA 0 stores the needles and then the sequence number into every rank's mailbox over NVLink; every CTA of every rank
// waits (bounded) for its own mailbox's sequence number and builds its needle tile from the mailbox; the last CTA to
// finish stores this rank's match count and then the sequence number into the root's mailbox, and on the root also
// waits for every rank's answer and writes the result to mapped host memory.  No other launch, no collective.
template <class Cfg, class Hook>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
match_rows_m_kernel(DbPlanes rows, uint64_t n_rows, uint32_t q, Tc4Sink sink, const __grid_constant__ Hook hook)
{
    constexpr int N = Cfg::N, T = Cfg::T, G = Cfg::G, S = Cfg::S, R = Cfg::R;
    constexpr bool MB = std::is_same<Hook, MailboxHook>::value;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sB = smem;                                                 // needle tile: N rows x 64 B, canonical K-major
    uint8_t *sR = smem + Cfg::NB_BYTES + 256;                           // cp.async ring: one 16-byte cell per (slot, tile, thread)
    uint64_t *bars = reinterpret_cast<uint64_t *>(sR + Cfg::RING_BYTES);
    uint64_t *a_full = bars, *d_full = bars + G * S;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d_full + G * S);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t n_units = (n_rows + Cfg::UNIT_ROWS - 1) / Cfg::UNIT_ROWS;

    if constexpr (MB) {
        if (hook.rank == hook.root && blockIdx.x == 0) {
            if (threadIdx.x < hook.n_vec) {
                const uint4 v = hook.needles_dev ? hook.needles_dev[threadIdx.x] : hook.needles.v[threadIdx.x];
                for (int r = 0; r < hook.world; r++)
                    hook.peers.p[r]->needles[threadIdx.x] = v;
            }
            __threadfence_system();
            __syncthreads();
            if ((int)threadIdx.x < hook.world)
                st_release_sys(&hook.peers.p[threadIdx.x]->seq_in, hook.seq);
        }
        if (threadIdx.x == 0) {
            const unsigned long long t0 = global_ns();
            while (ld_acquire_sys(&hook.mine->seq_in) != hook.seq) {
                if (global_ns() - t0 > kMailboxTimeoutNs) {
                    atomicExch(&hook.mine->error, 1u);
                    break;
                }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < N) {
        // needle j -> row j of the B tile (needle32: sign and inverse scale of the masked row encoding); its six
        // constant positions carry threshold - popc122(needle), so that D = threshold - hamming122
        const uint32_t j = threadIdx.x;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (j < q) {
            if constexpr (MB)      // written by a peer while this kernel was already running: not through the read-only path
                v = __ldcv(sink.needles_raw + 2 * (size_t)j);
            else
                v = __ldg(sink.needles_raw + 2 * (size_t)j);
        }
        const int32_t p = __popc(v.x & ~X1_BIAS_HASH_BITS) + __popc(v.y) + __popc(v.z) + __popc(v.w);
        const int32_t bias = j < q ? (int32_t)sink.threshold - p : -180;      // padding needles never flag
        uint4 c0 = needle32(v.x);
        c0.x = (c0.x & 0xFF000000u) | x1_bias_codes(bias);
        uint8_t *grp = sB + (j >> 3) * 512 + (j & 7u) * 16;
        *reinterpret_cast<uint4 *>(grp) = c0;
        *reinterpret_cast<uint4 *>(grp + 128) = needle32(v.y);
        *reinterpret_cast<uint4 *>(grp + 256) = needle32(v.z);
        *reinterpret_cast<uint4 *>(grp + 384) = needle32(v.w);
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < G * S; i++) {
            bar_init(&a_full[i], 4);
            bar_init(&d_full[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == Cfg::WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the needle tile is read by the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (warp < 4) {
        // scale factors 1.0 in all 128 lanes of columns 496..511
        const uint32_t base = tmem + (((uint32_t)warp * 32u) << 16) + X1_SF;
        const uint32_t one = 0x7F7F7F7Fu;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(base),
                     "r"(one)
                     : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (warp >= Cfg::WARP_MMA) {
        // ===== MMA issuers, one warp per group: per unit T x two MMAs (K = 64 each), A = a tile's 16 TMEM columns, B =
        // the needle tile.  One thread issues a small-N MMA every ~60 clk and a commit in ~54 (tools/x1_test.cu) -- a
        // latency of the issuing thread, not of the tensor pipe: four issuers reach four times the rate of one.
        if (lane == 0) {
            const uint32_t g = (uint32_t)warp - Cfg::WARP_MMA;
            const uint32_t n_it = x1_iterations<Cfg>(n_units, g);
            const uint64_t b0 = x1_desc(s32(sB), 128, 512), b1 = x1_desc(s32(sB) + 256, 128, 512);
            const uint32_t sf = tmem + X1_SF;
            auto mma = [&](uint32_t d, uint32_t a, uint64_t db, uint32_t acc) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%5], [%6], p;\n\t}" ::"r"(d),
                    "r"(a), "l"(db), "r"(Cfg::IDESC), "r"(acc), "r"(sf), "r"(sf + 8)
                    : "memory");
            };
            const uint32_t afull = s32(&a_full[g * S]);
            uint32_t s = 0, par = 0;
            for (uint32_t i = 0; i < n_it; i++) {
                // the issuer polls with a pause: its spin loop otherwise takes ~15 % of the SM's issue slots from the workers
                while (!bar_try_addr_nohint(afull + 8 * s, par))
                    asm volatile("nanosleep.u32 %0;" ::"n"(X1_ISSUER_SLEEP_NS));
                if (g == 0) X1_TR(1, i);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t col = tmem + (g * S + s) * Cfg::SLOT_COLS;
#pragma unroll
                for (int t = 0; t < T; t++) {
                    mma(col + 16 * T + N * t, col + 16 * t, b0, 0u);
                    mma(col + 16 * T + N * t, col + 16 * t + 8, b1, 1u);
                }
                umma_commit(&d_full[g * S + s]);
                if (g == 0) X1_TR(2, i);
                if (++s == (uint32_t)S) {
                    s = 0;
                    par ^= 1u;
                }
            }
        }
    } else {
        // ===== workers: T rows per thread per unit (row t * 128 + my of the unit) =====
        const uint32_t g = (uint32_t)warp >> 2, quad = (uint32_t)warp & 3u;
        const uint32_t n_it = x1_iterations<Cfg>(n_units, g);
        const uint32_t lane_tmem = tmem + ((quad * 32u) << 16) + g * S * Cfg::SLOT_COLS;
        const uint32_t my = quad * 32u + (uint32_t)lane;            // row within a tile
        const uint32_t afull = s32(&a_full[g * S]), dfull = s32(&d_full[g * S]);
        // this thread's ring cells: (ring slot sl, tile t) at cell + (sl * T + t) * CELL_STRIDE
        const uint32_t cell = s32(sR) + (g * X1_M + my) * 16u;
        constexpr uint32_t CELL_STRIDE = (uint32_t)G * X1_M * 16u;
        auto unit_row0 = [&](uint32_t i) { return (((uint64_t)i * G + g) * gridDim.x + blockIdx.x) * Cfg::UNIT_ROWS + my; };
        // unit i's copies: T x 16 bytes, one commit group per unit.  fetch() is called for i = 0, 1, 2 .. in order, so the
        // source is a running pointer; only a group's last unit can be partial (rows that do not exist are zero-filled:
        // src-size 0), every other unit is copied without a bounds check.
        const uint32_t n_it_full = x1_iterations<Cfg>(n_rows / Cfg::UNIT_ROWS, g);
        const uint4 *fp = rows.pre + unit_row0(0);
        const uint64_t fstride = (uint64_t)G * gridDim.x * Cfg::UNIT_ROWS;
        auto fetch = [&](uint32_t i, uint32_t slot) {
            const uint32_t dst = cell + slot * (T * CELL_STRIDE);
            if (i < n_it_full) {
#pragma unroll
                for (int t = 0; t < T; t++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + t * CELL_STRIDE), "l"(fp + t * X1_M) : "memory");
            } else if (i < n_it) {
                const uint64_t row0 = unit_row0(i);
#pragma unroll
                for (int t = 0; t < T; t++) {
                    const bool ok = row0 + (uint64_t)t * X1_M < n_rows;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + t * CELL_STRIDE),
                                 "l"(ok ? fp + t * X1_M : rows.pre), "r"(ok ? 16u : 0u)
                                 : "memory");
                }
            }
            fp += fstride;
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll 1
        for (uint32_t p = 0; p + 1 < (uint32_t)R; p++)
            fetch(p, p);
        static_assert((R & (R - 1)) == 0, "ring depth: a power of two");
        uint32_t sl = 0;       // ring slot of the unit being produced
        // Iteration i CONSUMES unit i - S (frees accumulator slot i % S) and PRODUCES unit i into the same slot.  The two are
        // interleaved so that the TMEM reads of the one and the TMEM writes of the other are in flight under the other's
        // integer work: [wait d_full, issue tcgen05.ld] [rows -> 16 operand words each, issue tcgen05.st] [wait::ld,
        // compare] [wait::st, hand the slot to the MMA warp].  CONSUME / PRODUCE are compile-time flags: the steady-state
        // loop below is unrolled over the S slots and carries no per-iteration conditions.
        auto step = [&](auto CONSUME, auto PRODUCE, uint32_t i, uint32_t s, uint32_t par_d) {
            constexpr bool consume = decltype(CONSUME)::value, produce = decltype(PRODUCE)::value;
            const uint32_t col = lane_tmem + s * Cfg::SLOT_COLS;
            uint32_t v[T][N];
            if constexpr (consume) {
                x1_wait(dfull + 8 * s, par_d);
                if (warp == 0 && lane == 0) X1_TR(3, i - S);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int t = 0; t < T; t++) {
                    if constexpr (N == 8)
                        T4_LD_X8(v[t], 0, col + 16 * T + N * t);
                    else if constexpr (N == 16)
                        T4_LD_X16(v[t], 0, col + 16 * T + N * t);
                    else if constexpr (N == 32)
                        T4_LD_X32(v[t], 0, col + 16 * T + N * t);
                    else
                        T4_LD_X64(v[t], 0, col + 16 * T + N * t);
                }
            }
            if constexpr (produce) {
                if (warp == 0 && lane == 0) X1_TR(5, i);
                fetch(i + R - 1, (sl - 1u) & (uint32_t)(R - 1));                   // refill the slot consumed one iteration ago
                asm volatile("cp.async.wait_group %0;" ::"n"(R - 1) : "memory");   // unit i's copies have landed
                if (warp == 0 && lane == 0) X1_TR(6, i);
                const uint32_t src = cell + sl * (T * CELL_STRIDE);
#pragma unroll
                for (int t = 0; t < T; t++) {
                    uint4 r;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                                 : "r"(src + t * CELL_STRIDE)
                                 : "memory");
                    uint4 c0 = mask32(r.x);
                    c0.x = (c0.x & 0xFF000000u) | X1_ROW_BIAS;
                    const uint4 c1 = mask32(r.y), c2 = mask32(r.z), c3 = mask32(r.w);
                    asm volatile(
                        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
                            col + 16 * t),
                        "r"(c0.x), "r"(c0.y), "r"(c0.z), "r"(c0.w), "r"(c1.x), "r"(c1.y), "r"(c1.z), "r"(c1.w), "r"(c2.x), "r"(c2.y),
                        "r"(c2.z), "r"(c2.w), "r"(c3.x), "r"(c3.y), "r"(c3.z), "r"(c3.w)
                        : "memory");
                }
                sl = (sl + 1u) & (uint32_t)(R - 1);
            }
            if constexpr (consume) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (warp == 0 && lane == 0) X1_TR(4, i - S);
                uint32_t below = 0xFFFFFFFFu;       // sign bit survives iff every accumulator of every row is negative
#pragma unroll
                for (int t = 0; t < T; t++)
                    below &= x1_and_tree<N>(v[t]);
                if (__any_sync(0xFFFFFFFFu, (int32_t)below >= 0)) {
                    if ((int32_t)below >= 0) {
#pragma unroll
                        for (int t = 0; t < T; t++) {
                            const uint64_t row = unit_row0(i - S) + (uint64_t)t * X1_M;
                            if (row < n_rows) {
#pragma unroll
                                for (int j = 0; j < N; j++)
                                    if ((int32_t)v[t][j] >= 0)     // padding needles (j >= q) are always negative
                                        x1_verify(row, (uint32_t)j, rows, sink);
                            }
                        }
                    }
                }
            }
            if constexpr (produce) {
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (lane == 0)
                    bar_arrive_addr(afull + 8 * s);
                if (warp == 0 && lane == 0) X1_TR(7, i);
            }
        };
        constexpr std::true_type yes{};
        constexpr std::false_type no{};
        // first round: nothing to consume yet
#pragma unroll
        for (uint32_t s = 0; s < (uint32_t)S; s++)
            if (s < n_it)
                step(no, yes, s, s, 0u);
        // steady state: rounds in which every slot both consumes and produces
        const uint32_t full_rounds = n_it / S;
        for (uint32_t r = 1; r < full_rounds; r++) {
#pragma unroll
            for (uint32_t s = 0; s < (uint32_t)S; s++)
                step(yes, yes, r * S + s, s, (r - 1u) & 1u);
        }
        // last units and the drain
#pragma unroll 1
        for (uint32_t i = max((uint32_t)S, full_rounds * S); i < n_it + S; i++) {
            const uint32_t s = i % S, par_d = (i / S - 1u) & 1u;
            if (i < n_it)
                step(yes, yes, i, s, par_d);
            else
                step(yes, no, i, s, par_d);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == Cfg::WARP_MMA)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    if constexpr (MB) {
        if (warp == 0) {
            unsigned int last = 0;
            if (lane == 0) {
                __threadfence();
                last = atomicAdd(&hook.mine->done, 1u) == gridDim.x - 1u;
            }
            last = __shfl_sync(0xFFFFFFFFu, last, 0);
            if (last) {
                MailboxMem *root = hook.peers.p[hook.root];
                if (lane == 0) {
                    __threadfence();
                    hook.mine->done = 0;
                    const unsigned int cnt = atomicAdd(&hook.mine->count[hook.seq & 1u], 0u);
                    const unsigned int err = atomicAdd(&hook.mine->error, 0u);
                    root->result[hook.rank] = (cnt & 0x7FFFFFFFu) | (err ? 0x80000000u : 0u);
                    hook.mine->count[(hook.seq + 1u) & 1u] = 0;
                    __threadfence_system();
                    st_release_sys(&root->seq_out[hook.rank], hook.seq);
                }
                __syncwarp();
                if (hook.rank == hook.root)
                    mailbox_collect(hook.mine, hook.world, hook.seq, hook.host_out);
            }
        }
    }
}

template <class Cfg, class Hook = NoHook>
int launch_x1(DbPlanes rows, size_t n_rows, uint32_t q, const Tc4Sink &sink, int sm_count, cudaStream_t s, const Hook &hook = Hook())
{
    const uint64_t n_units = (n_rows + Cfg::UNIT_ROWS - 1) / Cfg::UNIT_ROWS;
    const unsigned grid = (unsigned)std::min<uint64_t>(n_units, (uint64_t)sm_count);
    YPDQ_CUDA_TRY(
        cudaFuncSetAttribute(match_rows_m_kernel<Cfg, Hook>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    match_rows_m_kernel<Cfg, Hook><<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(rows, n_rows, q, sink, hook);
    YPDQ_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// 1 .. 64 needles, any threshold up to 128 (the prefix filter only thins the pairs; it is useful up to ~48)
int launch_match_tensor_rows(DbPlanes rows, size_t n_rows, uint64_t row_offset, const uint8_t *needles, size_t q,
                             uint32_t threshold, YumePDQ_Match *out, uint32_t cap, uint32_t *count,
                             unsigned long long *min_key, int mode, int sm_count, cudaStream_t s, NeedleMap needle_base)
{
    if (n_rows == 0 || q == 0)
        return 0;
    if (q > 64 || threshold > 128)
        return YUME_PDQ_CUDA_E_INVALID;
    Tc4Sink sink{out, cap, count, min_key, row_offset, (uint64_t)q, mode, needle_base, reinterpret_cast<const uint4 *>(needles), threshold};
    if (q <= 8)
        return launch_x1<X1N8>(rows, n_rows, (uint32_t)q, sink, sm_count, s);
    if (q <= 16)
        return launch_x1<X1N16>(rows, n_rows, (uint32_t)q, sink, sm_count, s);
    if (q <= 32)
        return launch_x1<X1N32>(rows, n_rows, (uint32_t)q, sink, sm_count, s);
    return launch_x1<X1N64>(rows, n_rows, (uint32_t)q, sink, sm_count, s);
}

// The sharded query in one launch per rank (1 .. 32 needles, n_rows > 0): counts this shard's matches into the mailbox and
// carries the exchange with the peers (see match_rows_m_kernel).  Every CTA spins on the mailbox, so the grid must be
// resident at once: it is at most one CTA per SM.
int launch_match_tensor_rows_mailbox(DbPlanes rows, size_t n_rows, uint64_t row_offset, size_t q, uint32_t threshold,
                                     const MailboxHook &hook, int sm_count, cudaStream_t s)
{
    if (n_rows == 0 || q == 0 || q > (size_t)kMailboxMaxNeedles || threshold > 128)
        return YUME_PDQ_CUDA_E_INVALID;
    Tc4Sink sink{nullptr, 0, &hook.mine->count[hook.seq & 1u], nullptr, row_offset, (uint64_t)q, MODE_LIST, 0,
                 reinterpret_cast<const uint4 *>(hook.mine->needles), threshold};
    if (q <= 8)
        return launch_x1<X1N8>(rows, n_rows, (uint32_t)q, sink, sm_count, s, hook);
    if (q <= 16)
        return launch_x1<X1N16>(rows, n_rows, (uint32_t)q, sink, sm_count, s, hook);
    return launch_x1<X1N32>(rows, n_rows, (uint32_t)q, sink, sm_count, s, hook);
}

}  // namespace ypdq
