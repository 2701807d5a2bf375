// matcher_tc.cu -- tensor-core (tcgen05) formulation of the Hamming matcher for LARGE needle batches.
//
// Same contract as matcher.cu (reference src/matching/mod.rs:336-362: every (row, needle) with
// popcount(row ^ needle) <= threshold, bit-exact), different arithmetic: with bits mapped to +-1 int8,
//     dot(needle, row) = 256 - 2 * hamming(needle, row),
// so a 128-needle x 256-row x K=256 tile is one int8 GEMM tile on the 5th-generation tensor cores
// (tcgen05.mma kind::i8, s32 accumulators in TMEM; exact).  The POPC-pipe kernels of matcher.cu cap at
// ~2.2e12 pairs/s (XU pipe 97 % busy); this path moves the pair arithmetic to the tensor pipe.
// (mma.sync b1 is emulated on sm_100a, so +-1 int8 is the native formulation.)
//
// Persistent kernel, one CTA per SM, 10 warps, warp-specialised:
//   warp 0      A producer: streams pre-expanded needle tiles (32 KB, canonical K-major SWIZZLE_NONE
//               layout) from global/L2 into a 3-stage shared-memory ring with TMA bulk copies
//   warp 1      MMA issuer: one elected lane issues 8 x tcgen05.mma (K = 32 each) per tile into one of
//               two 256-column TMEM accumulator stages, tcgen05.commit -> mbarriers
//   warps 2-5   epilogue: tcgen05.ld 32x32b (lane = needle, column = row), 3-input max tree, rare hit path
//   warps 6-9   B expanders: turn this CTA's next 256 database rows (8 KB of bits) into the 64 KB +-1 tile
//               (double buffered) -- done once per row tile and reused by every needle tile
// Shared memory: 2 x 64 KB (B) + 3 x 32 KB (A) = 224 KB.  TMEM: all 512 columns (2 stages).
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace ypdq {
namespace {

constexpr int TC_M = 128;                 // needles per tile (TMEM lanes)
constexpr int TC_N = 256;                 // rows per tile (TMEM columns)
constexpr int TC_K = 256;                 // bits
constexpr int TC_A_BYTES = TC_M * TC_K;   // 32 KB
constexpr int TC_B_BYTES = TC_N * TC_K;   // 64 KB
constexpr int TC_A_STAGES = 3;
constexpr int TC_THREADS = 320;
constexpr uint32_t TC_LBO = 128;          // bytes between the two 16-byte K chunks of one MMA
constexpr uint32_t TC_SBO = 2048;         // bytes between 8-row groups
constexpr size_t TC_SMEM = 2 * (size_t)TC_B_BYTES + TC_A_STAGES * (size_t)TC_A_BYTES + 256;

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t *b, uint32_t n)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n));
}
__device__ __forceinline__ void bar_arrive(uint64_t *b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint64_t *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t *b, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(s32(b)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
                 "l"(src), "r"(bytes), "r"(s32(b))
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *b)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr)
{
    // K-major, SWIZZLE_NONE canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units; version 1
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (uint64_t)(TC_LBO >> 4) << 16 | (uint64_t)(TC_SBO >> 4) << 32 |
           (uint64_t)1 << 46;
}
// instruction descriptor: D = s32, A = B = signed int8, both K-major, N = 256, M = 128
constexpr uint32_t TC_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) |
                              ((uint32_t)(TC_M >> 4) << 24);

// 4 bits -> 4 bytes of +-1 (bit 1 -> 0x01, bit 0 -> 0xFF)
__device__ __forceinline__ uint32_t expand4(uint32_t nib)
{
    uint32_t m = (nib * 0x00204081u) & 0x01010101u;
    return ~(m * 0xFEu);
}
// 16 bits -> 16 bytes, K order = bit order (bit i of byte b of the hash is K index 8b + i)
__device__ __forceinline__ uint4 expand16(uint32_t h)
{
    return make_uint4(expand4(h & 15u), expand4((h >> 4) & 15u), expand4((h >> 8) & 15u), expand4((h >> 12) & 15u));
}
// canonical tile offset of (row r, 16-byte chunk kc)
__device__ __forceinline__ uint32_t tile_off(uint32_t r, uint32_t kc)
{
    return (r >> 3) * TC_SBO + kc * TC_LBO + (r & 7u) * 16u;
}

// needles (q x 32 B) -> MT tiles of 32 KB in the canonical layout; padding needles expand from zero bits
__global__ void __launch_bounds__(256)
expand_needles_kernel(const uint4 *__restrict__ needles, uint32_t q, uint8_t *__restrict__ out, uint32_t n_tiles)
{
    const uint32_t idx = blockIdx.x * 256 + threadIdx.x;  // one thread per (needle, half)
    const uint32_t j = idx >> 1, half = idx & 1u;
    if (j >= n_tiles * TC_M)
        return;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (j < q)
        v = __ldg(needles + 2 * (size_t)j + half);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint8_t *tile = out + (size_t)(j / TC_M) * TC_A_BYTES;
    const uint32_t r = j % TC_M;
#pragma unroll
    for (int c = 0; c < 8; c++) {  // this half covers chunks 8*half .. 8*half+7 (16 bits each)
        const uint32_t h = (w[c >> 1] >> ((c & 1) * 16)) & 0xFFFFu;
        *reinterpret_cast<uint4 *>(tile + tile_off(r, 8 * half + c)) = expand16(h);
    }
}

struct TcSink {
    YumePDQ_Match *out;
    uint32_t cap;
    uint32_t *count;
    unsigned long long *min_key;
    uint64_t row_offset;
    uint64_t q;
    int mode;
    NeedleMap needle_base;   // added to reported needle indices (LIST mode): pipelined callers match in chunks
};

__device__ __forceinline__ void tc_emit(const TcSink &k, uint64_t local_row, uint32_t needle, uint32_t dist)
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

#define TMEM_LD_X32(v, taddr)                                                                                  \
    asm volatile(                                                                                              \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                              \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                             \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"              \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),      \
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),          \
          "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),          \
          "=r"(v[30]), "=r"(v[31])                                                                             \
        : "r"(taddr))

__global__ void __launch_bounds__(TC_THREADS, 1)
match_tensor_kernel(DbPlanes rows, uint64_t n_rows, const uint8_t *__restrict__ needles_exp,
                    uint32_t n_mtiles, uint32_t q, int32_t dot_min, TcSink sink)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sB = smem;                                   // 2 x 64 KB
    uint8_t *sA = smem + 2 * TC_B_BYTES;                  // 3 x 32 KB
    uint64_t *bars = reinterpret_cast<uint64_t *>(sA + TC_A_STAGES * TC_A_BYTES);
    uint64_t *a_full = bars, *a_empty = bars + 3, *b_full = bars + 6, *b_empty = bars + 8;
    uint64_t *d_full = bars + 10, *d_empty = bars + 12;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 14);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_A_STAGES; i++) {
            bar_init(&a_full[i], 1);
            bar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            bar_init(&b_full[i], 128);
            bar_init(&b_empty[i], 1);
            bar_init(&d_full[i], 1);
            bar_init(&d_empty[i], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    const uint64_t n_ntiles = (n_rows + TC_N - 1) / TC_N;

    if (warp == 0) {
        // ===== A producer =====
        if (lane == 0) {
            uint32_t it = 0;
            for (uint64_t nt = blockIdx.x; nt < n_ntiles; nt += gridDim.x)
                for (uint32_t m = 0; m < n_mtiles; m++, it++) {
                    const uint32_t s = it % TC_A_STAGES, use = it / TC_A_STAGES;
                    bar_wait(&a_empty[s], (use & 1u) ^ 1u);
                    bar_expect_tx(&a_full[s], TC_A_BYTES);
                    bulk_g2s(sA + (size_t)s * TC_A_BYTES, needles_exp + (size_t)m * TC_A_BYTES, TC_A_BYTES, &a_full[s]);
                }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t it = 0, i = 0;
            for (uint64_t nt = blockIdx.x; nt < n_ntiles; nt += gridDim.x, i++) {
                const uint32_t b = i & 1u;
                bar_wait(&b_full[b], (i >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_addr = s32(sB + (size_t)b * TC_B_BYTES);
                for (uint32_t m = 0; m < n_mtiles; m++, it++) {
                    const uint32_t s = it % TC_A_STAGES, acc = it & 1u;
                    bar_wait(&a_full[s], (it / TC_A_STAGES) & 1u);
                    bar_wait(&d_empty[acc], ((it >> 1) & 1u) ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = s32(sA + (size_t)s * TC_A_BYTES);
                    const uint32_t d_addr = tmem + acc * TC_N;
#pragma unroll
                    for (uint32_t k = 0; k < TC_K / 32; k++) {
                        const uint64_t da = smem_desc(a_addr + k * 2 * TC_LBO);
                        const uint64_t db = smem_desc(b_addr + k * 2 * TC_LBO);
                        const uint32_t accumulate = k > 0 ? 1u : 0u;
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t"
                            "setp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_addr),
                            "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate)
                            : "memory");
                    }
                    umma_commit(&a_empty[s]);     // A stage may be refilled once these MMAs retire
                    umma_commit(&d_full[acc]);    // accumulator ready for the epilogue
                }
                umma_commit(&b_empty[b]);         // B buffer may be overwritten
            }
        }
    } else if (warp < 6) {
        // ===== epilogue: TMEM lane quadrant = warp % 4 =====
        const uint32_t quad = warp & 3u;
        uint32_t it = 0;
        for (uint64_t nt = blockIdx.x; nt < n_ntiles; nt += gridDim.x)
            for (uint32_t m = 0; m < n_mtiles; m++, it++) {
                const uint32_t acc = it & 1u;
                bar_wait(&d_full[acc], (it >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t needle = m * TC_M + quad * 32 + lane;
                const uint32_t tbase = tmem + ((quad * 32u) << 16) + acc * TC_N;
#pragma unroll 1
                for (uint32_t c0 = 0; c0 < TC_N; c0 += 32) {
                    int32_t v[32];
                    TMEM_LD_X32(v, tbase + c0);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    int32_t mx = v[0];
#pragma unroll
                    for (int j = 1; j < 32; j++)
                        mx = max(mx, v[j]);
                    if (mx >= dot_min && needle < q) {  // rare: some row of this chunk matches this needle
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            if (v[j] >= dot_min) {
                                const uint64_t row = nt * TC_N + c0 + j;
                                if (row < n_rows)
                                    tc_emit(sink, row, needle, (uint32_t)(256 - v[j]) >> 1);
                            }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0)
                    bar_arrive(&d_empty[acc]);
            }
    } else {
        // ===== B expanders: 128 threads, thread t expands rows t and t + 128 of the tile =====
        const uint32_t t = threadIdx.x - 192;
        uint32_t i = 0;
        for (uint64_t nt = blockIdx.x; nt < n_ntiles; nt += gridDim.x, i++) {
            const uint32_t b = i & 1u;
            bar_wait(&b_empty[b], ((i >> 1) & 1u) ^ 1u);
            uint8_t *tile = sB + (size_t)b * TC_B_BYTES;
#pragma unroll
            for (uint32_t h = 0; h < 2; h++) {
                const uint32_t r = t + 128 * h;
                const uint64_t row = nt * TC_N + r;
                uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
                if (row < n_rows) {
                    lo = __ldg(rows.pre + row);
                    hi = __ldg(rows.suf + row);
                }
                const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                for (uint32_t kc = 0; kc < 16; kc++) {
                    const uint32_t hbits = (w[kc >> 1] >> ((kc & 1u) * 16)) & 0xFFFFu;
                    *reinterpret_cast<uint4 *>(tile + tile_off(r, kc)) = expand16(hbits);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
            bar_arrive(&b_full[b]);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace

int launch_match_tensor(DbPlanes rows, size_t n_rows, uint64_t row_offset, const uint8_t *needles, size_t q,
                        uint32_t threshold, YumePDQ_Match *out, uint32_t cap, uint32_t *count,
                        unsigned long long *min_key, int mode, int sm_count, cudaStream_t s, NeedleMap needle_base)
{
    if (n_rows == 0 || q == 0)
        return 0;
    const uint32_t n_mtiles = (uint32_t)((q + TC_M - 1) / TC_M);
    uint8_t *exp = nullptr;
    cudaError_t e = cudaMallocAsync(&exp, (size_t)n_mtiles * TC_A_BYTES, s);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? YUME_PDQ_CUDA_E_NOMEM : (int)e;
    }
    const uint32_t threads = n_mtiles * TC_M * 2;
    expand_needles_kernel<<<(threads + 255) / 256, 256, 0, s>>>(reinterpret_cast<const uint4 *>(needles), (uint32_t)q, exp,
                                                                n_mtiles);
    YPDQ_LAUNCH_CHECK();
    // per launch: function attributes are per device, and the call costs microseconds
    YPDQ_CUDA_TRY(cudaFuncSetAttribute(match_tensor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    TcSink sink{out, cap, count, min_key, row_offset, (uint64_t)q, mode, needle_base};
    const uint64_t n_ntiles = (n_rows + TC_N - 1) / TC_N;
    const unsigned grid = (unsigned)std::min<uint64_t>(n_ntiles, (uint64_t)sm_count);
    const int32_t dot_min = 256 - 2 * (int32_t)std::min<uint32_t>(threshold, 256u);
    match_tensor_kernel<<<grid, TC_THREADS, TC_SMEM, s>>>(rows, n_rows, exp, n_mtiles,
                                                          (uint32_t)q, dot_min, sink);
    YPDQ_LAUNCH_CHECK();
    YPDQ_CUDA_TRY(cudaFreeAsync(exp, s));
    return 0;
}

}  // namespace ypdq
