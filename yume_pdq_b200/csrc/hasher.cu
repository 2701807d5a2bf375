// hasher.cu -- batched PDQ-variant hasher for sm_100a.
//
// One CTA of 128 threads hashes one 512x512 f32 frame in a single pass over the
// frame and emits all 8 dihedral hashes.  The arithmetic is the reference scalar
// kernel's (DefaultKernel, /root/reference src/kernel/mod.rs), operation by operation, for finite
// input; two conventions are NOT the scalar kernel's and are listed in DESIGN.md section 5: the
// dihedral sign flips follow the AVX2 kernel (pdqf_negate_alt_rows::<true> negates EVEN rows,
// x86.rs:285-326; the scalar trait default negates odd rows, mod.rs:392-405, so only NORMAL and
// FLOPPED agree with DefaultKernel bit for bit), and the zero / ~1e-40 border taps of the 7x7 tent
// LUT are not multiplied in (a NaN or Inf pixel in the border ring propagates in the reference only):
//   jarosz_compress   mod.rs:644-666   tent 7x7 (5x5 effective), di outer / dj inner,
//                                      separate f32 multiply and add (no FMA)
//   dct2d             mod.rs:543-574   D(16x127) B D^T with 4 round-robin partial
//                                      sums per dot product, ((s0+s1)+s2)+s3
//   sum_of_gradients  mod.rs:617-642   sequential f32 accumulation
//   quantize          mod.rs:595-608   Torben lower median sorted[127], bit = v > median
//   visit_dihedrals   lib.rs:358-420   sign flips / transpose in frequency space
//                                      (even-row convention of x86.rs:285-326)
// so __fmul_rn/__fadd_rn are used everywhere a rounding happens (never
// contracted to FMA) and the summation orders are the reference's.
//
// Mapping: thread j owns output column j of the 127x127 tent image.
// For every tent row oi it loads its 5x5 window straight from global memory with
// 128-bit loads (a warp reads 512 contiguous bytes per instruction; the row
// shared between consecutive windows stays in registers), produces B[oi][j] and
// immediately applies the rank-1 update T[k][j] += D[k][oi] * B[oi][j] for the 16
// DCT rows -- 64 accumulators (16 k x 4 partial sums) in registers, so the
// 127x127 intermediate never exists in memory.  No synchronisation happens while
// the frame streams; one lane prefetches the rows 8 windows ahead into L2
// (cp.async.bulk.prefetch.L2).  The two ~1e-40 residue weights of the reference
// LUT (TENT[0][4], TENT[0][5]) are not applied: they cannot change any result
// unless a whole 5x5 window is exactly zero (documented in DESIGN.md).
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <type_traits>

#include "../../include/yume_pdq_luts.h"
#include "common.cuh"

namespace ypdq {

namespace {

constexpr int HT = 128;          // threads per CTA
constexpr int TSTRIDE = 129;     // padded row stride of the 16x127 stage-1 result
// L2 prefetch distance in tent rows (4 input rows = 8 KB of f32 each).  Measured on a B200, 4096 f32 frames:
// 2 -> 696 us, 3 -> 707, 4 -> 673, 5 -> 668, 6 -> 660 (6.50 TB/s), 7 -> 679, 8 -> 756, 16 -> 928, 32 -> 1168: beyond
// 6 windows x 592 resident CTAs the prefetched lines start to be evicted from L2 before they are used (ncu at 8:
// DRAM reads 1.22 x the frame bytes).
#ifndef HASH_PREFETCH_AHEAD_CFG
#define HASH_PREFETCH_AHEAD_CFG 6
#endif
constexpr int PREFETCH_AHEAD = HASH_PREFETCH_AHEAD_CFG;

__device__ __align__(16) float g_dt[127 * 16];          // D transposed, g_dt[k2*16 + k] = D[k][k2]
__constant__ uint32_t c_tent_bits[49];    // W[di][dj], read as constant-bank operands
__constant__ float c_luma[3] = {0.299f, 0.587f, 0.114f};  // RGB8_TO_LUMA8_TABLE_ITU, constants.rs:23

struct __align__(16) HashSmem {
    float dt[127 * 16];        // D transposed: dt[k2*16 + k] = D[k][k2]
    float t[16 * TSTRIDE];     // stage-1 result T[k][j]
    float p[256];              // PDQF
    float thr[4];              // medians of P, C.P, R.C.P, R.P
};

__device__ __forceinline__ void prefetch_l2(const void *p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// weight W[di][dj] of the tent LUT as a compile-time-indexed constant load
__device__ __forceinline__ float tent_w(int di, int dj)
{
    return __uint_as_float(c_tent_bits[di * 7 + dj]);
}

// the five f32 luma taps of one window row: input columns 4j+2 .. 4j+6
struct Win5 {
    float v0, v1, v2, v3, v4;
};

// ---- window row DI (1..5): five taps in reference order (dj = 1..5) ----
template <int DI>
__device__ __forceinline__ float tent_row(float sum, const Win5 &w)
{
    sum = __fadd_rn(sum, __fmul_rn(tent_w(DI, 1), w.v0));
    sum = __fadd_rn(sum, __fmul_rn(tent_w(DI, 2), w.v1));
    sum = __fadd_rn(sum, __fmul_rn(tent_w(DI, 3), w.v2));
    sum = __fadd_rn(sum, __fmul_rn(tent_w(DI, 4), w.v3));
    sum = __fadd_rn(sum, __fmul_rn(tent_w(DI, 5), w.v4));
    return sum;
}

// ---- pixel formats: the conversion the reference CLI applies to every input row
// before hash_get_threshold (main.rs:486-608) is fused into the window load, so a u8 frame is read
// from HBM once, as bytes.
//   LUMA8  `row_buf[j] as f32`                                     main.rs:508-510
//   RGB8   Kernel::cvt_rgb8_to_luma8f,  cr*r + cg*g + cb*b         mod.rs:303-321
//   RGBA8  Kernel::cvt_rgba8_to_luma8f, alpha ignored              mod.rs:324-340
// FMA = false is the trait default (separate multiplies and adds, left to right); FMA = true is
// the rounding of the AVX2 override (mul, fmadd, fmadd; x86.rs:686-753).
__device__ __forceinline__ float u8f(uint32_t w, int byte)
{
    return (float)((w >> (8 * byte)) & 0xffu);   // exact
}

template <bool FMA>
__device__ __forceinline__ float luma_of(float r, float g, float b)
{
    if (FMA)
        return __fmaf_rn(b, c_luma[2], __fmaf_rn(g, c_luma[1], __fmul_rn(r, c_luma[0])));
    return __fadd_rn(__fadd_rn(__fmul_rn(c_luma[0], r), __fmul_rn(c_luma[1], g)), __fmul_rn(c_luma[2], b));
}

template <bool FMA>
__device__ __forceinline__ float luma_rgba(uint32_t px)   // little-endian R, G, B, A
{
    return luma_of<FMA>(u8f(px, 0), u8f(px, 1), u8f(px, 2));
}

template <int FMT, bool FMA>
struct Px;
template <bool FMA>
struct Px<YUME_PDQ_CUDA_PIXEL_F32, FMA> {
    static constexpr int BPP = 4;
    // p = row + 4j pixels; columns 4j..4j+7 in two 128-bit loads, the window is columns 4j+2..4j+6
    static __device__ __forceinline__ Win5 load(const uint8_t *p)
    {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
        return {a.z, a.w, b.x, b.y, b.z};
    }
};
template <bool FMA>
struct Px<YUME_PDQ_CUDA_PIXEL_LUMA8, FMA> {
    static constexpr int BPP = 1;
    static __device__ __forceinline__ Win5 load(const uint8_t *p)
    {
        const uint32_t a = __ldg(reinterpret_cast<const uint32_t *>(p));
        const uint32_t b = __ldg(reinterpret_cast<const uint32_t *>(p) + 1);
        return {u8f(a, 2), u8f(a, 3), u8f(b, 0), u8f(b, 1), u8f(b, 2)};
    }
};
template <bool FMA>
struct Px<YUME_PDQ_CUDA_PIXEL_RGB8, FMA> {
    static constexpr int BPP = 3;
    // pixels 4j+2..4j+6 are bytes 6..20 of p = row + 12j: words 1..5
    static __device__ __forceinline__ Win5 load(const uint8_t *p)
    {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(p);
        const uint32_t w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3), w4 = __ldg(q + 4),
                       w5 = __ldg(q + 5);
        return {luma_of<FMA>(u8f(w1, 2), u8f(w1, 3), u8f(w2, 0)),
                luma_of<FMA>(u8f(w2, 1), u8f(w2, 2), u8f(w2, 3)),
                luma_of<FMA>(u8f(w3, 0), u8f(w3, 1), u8f(w3, 2)),
                luma_of<FMA>(u8f(w3, 3), u8f(w4, 0), u8f(w4, 1)),
                luma_of<FMA>(u8f(w4, 2), u8f(w4, 3), u8f(w5, 0))};
    }
};
template <bool FMA>
struct Px<YUME_PDQ_CUDA_PIXEL_RGBA8, FMA> {
    static constexpr int BPP = 4;
    static __device__ __forceinline__ Win5 load(const uint8_t *p)
    {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(p) + 1);
        return {luma_rgba<FMA>(a.z), luma_rgba<FMA>(a.w), luma_rgba<FMA>(b.x), luma_rgba<FMA>(b.y),
                luma_rgba<FMA>(b.z)};
    }
};

// value of dihedral d's PDQF at (i, j), from the untransformed PDQF in shared memory.
// d: 0 P, 1 C.P, 2 R.C.P, 3 R.P, 4 P^T, 5 R.P^T, 6 C.R.P^T, 7 C.P^T with
// C = negate even columns (pdqf_negate_alt_cols::<false>, mod.rs:368-382) and
// R = negate even rows (pdqf_negate_alt_rows::<true>, x86.rs:285-326).
__device__ __forceinline__ float dihedral_value(const float *p, int d, int i, int j)
{
    float v = (d & 4) ? p[j * 16 + i] : p[i * 16 + j];
    const bool c = (0xC6 >> d) & 1;  // d in {1,2,6,7}
    const bool r = (0x6C >> d) & 1;  // d in {2,3,5,6}
    const bool neg = (c && !(j & 1)) != (r && !(i & 1));
    return neg ? -v : v;
}

// order-preserving map f32 -> u32 (for the radix select)
__device__ __forceinline__ uint32_t f2key(float f)
{
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Lower median sorted[127] of the 256 values of sign-variant `variant` (0 P, 1 C.P,
// 2 R.C.P, 3 R.P): what torben_median (mod.rs:212-253) returns.  One warp, MSB-first
// radix select over order-preserving keys; 8 values per lane.
__device__ float warp_lower_median(const float *p, int variant, int lane)
{
    uint32_t key[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        int idx = lane + 32 * e;
        key[e] = f2key(dihedral_value(p, variant, idx >> 4, idx & 15));
    }
    uint32_t alive = 0xffu, prefix = 0, k = 127;
    for (int bit = 31; bit >= 0; bit--) {
        uint32_t zeros = 0;
#pragma unroll
        for (int e = 0; e < 8; e++)
            zeros |= (((key[e] >> bit) & 1u) ^ 1u) << e;
        uint32_t cnt0 = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(alive & zeros));
        if (k < cnt0) {
            alive &= zeros;
        } else {
            k -= cnt0;
            alive &= ~zeros;
            prefix |= 1u << bit;
        }
    }
    return key2f(prefix);
}

// sum_of_gradients (mod.rs:617-642) of dihedral d, sequential like the reference
__device__ float gradient_sum(const float *p, int d)
{
    float g = 0.0f;
    for (int i = 0; i < 15; i++)
        for (int j = 0; j < 16; j++)
            g = __fadd_rn(g, fabsf(__fsub_rn(dihedral_value(p, d, i, j), dihedral_value(p, d, i + 1, j))));
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 15; j++)
            g = __fadd_rn(g, fabsf(__fsub_rn(dihedral_value(p, d, i, j), dihedral_value(p, d, i, j + 1))));
    return g;
}

__device__ __forceinline__ float adjust_quality(float g)
{
    float s = __fdiv_rn(g, 180.0f);  // QUALITY_ADJUST_DIVISOR, mod.rs:257
    return s < 1.0f ? s : 1.0f;      // f32::min(scaled, 1.0): NaN -> 1.0
}

// 256 hash bits of dihedral d against threshold thr; all 128 threads participate.
// Flat index i = 16*row + col -> byte 31 - i/8, bit i%8 (mod.rs:595-608, threshold.rs:37-119).
__device__ __forceinline__ void write_hash_bits(const float *p, int d, float thr, uint8_t *out,
                                                int tid)
{
    uint32_t *out32 = reinterpret_cast<uint32_t *>(out);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        int idx = tid + 128 * h;  // warp w covers ballot word b = w + 4h
        bool bit = dihedral_value(p, d, idx >> 4, idx & 15) > thr;
        uint32_t b = __ballot_sync(0xffffffffu, bit);
        if ((tid & 31) == 0) {
            int word = (tid >> 5) + 4 * h;                 // indices 32*word .. 32*word+31
            out32[7 - word] = __byte_perm(b, 0, 0x0123);   // byte 31-4*word holds bits 0..7
        }
    }
}

// Epilogue shared by the fused kernel and the per-stage entry points: from the PDQF in
// sm.p compute medians, 8 hashes, 8 qualities, 8 thresholds.
__device__ void dihedral_epilogue(HashSmem &sm, int tid, uint8_t *hashes, float *quality,
                                  float *thresholds, bool have_thr0, float thr0)
{
    const int warp = tid >> 5, lane = tid & 31;
    // visit_dihedrals (lib.rs:358-420): thr0 is the caller's threshold for NORMAL; the three
    // sign variants get a fresh quantize
    float med = 0.0f;
    if (!(have_thr0 && warp == 0))
        med = warp_lower_median(sm.p, warp, lane);
    if (lane == 0)
        sm.thr[warp] = (have_thr0 && warp == 0) ? thr0 : med;
    __syncthreads();
#pragma unroll 1
    for (int d = 0; d < 8; d++)
        write_hash_bits(sm.p, d, sm.thr[d & 3], hashes + 32 * d, tid);
    if (tid < 8) {
        if (thresholds)
            thresholds[tid] = sm.thr[tid & 3];
        if (quality)
            quality[tid] = adjust_quality(gradient_sum(sm.p, tid));
    }
}

__device__ __forceinline__ void load_dt(HashSmem &sm, int tid)
{
    const float4 *src = reinterpret_cast<const float4 *>(g_dt);
    float4 *dst = reinterpret_cast<float4 *>(sm.dt);
    for (int i = tid; i < 127 * 4; i += HT)
        dst[i] = src[i];
}

// stage 2 of dct2d (mod.rs:563-570): PDQF[k][jj] = sum_m T[k][m] * D[jj][m], 4 partial sums
__device__ __forceinline__ void dct_stage2(HashSmem &sm, int tid)
{
    const int k = tid >> 3, jj = (tid & 7) * 2;
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
    const float *trow = sm.t + k * TSTRIDE;
#pragma unroll 4
    for (int m = 0; m < 124; m += 4) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            float tv = trow[m + r];
            float2 dv = *reinterpret_cast<const float2 *>(&sm.dt[(m + r) * 16 + jj]);
            s0[r] = __fadd_rn(s0[r], __fmul_rn(tv, dv.x));
            s1[r] = __fadd_rn(s1[r], __fmul_rn(tv, dv.y));
        }
    }
#pragma unroll
    for (int r = 0; r < 3; r++) {
        float tv = trow[124 + r];
        float2 dv = *reinterpret_cast<const float2 *>(&sm.dt[(124 + r) * 16 + jj]);
        s0[r] = __fadd_rn(s0[r], __fmul_rn(tv, dv.x));
        s1[r] = __fadd_rn(s1[r], __fmul_rn(tv, dv.y));
    }
    sm.p[k * 16 + jj] = __fadd_rn(__fadd_rn(__fadd_rn(s0[0], s0[1]), s0[2]), s0[3]);
    sm.p[k * 16 + jj + 1] = __fadd_rn(__fadd_rn(__fadd_rn(s1[0], s1[1]), s1[2]), s1[3]);
}

// rank-1 update of the 16x4 accumulators with tent value b of row k2 (k2 & 3 == R)
template <int R>
__device__ __forceinline__ void rank1(float (&acc)[16][4], const float *dt_row, float b)
{
#pragma unroll
    for (int kq = 0; kq < 4; kq++) {
        const float4 d = *reinterpret_cast<const float4 *>(dt_row + 4 * kq);
        acc[4 * kq + 0][R] = __fadd_rn(acc[4 * kq + 0][R], __fmul_rn(d.x, b));
        acc[4 * kq + 1][R] = __fadd_rn(acc[4 * kq + 1][R], __fmul_rn(d.y, b));
        acc[4 * kq + 2][R] = __fadd_rn(acc[4 * kq + 2][R], __fmul_rn(d.z, b));
        acc[4 * kq + 3][R] = __fadd_rn(acc[4 * kq + 3][R], __fmul_rn(d.w, b));
    }
}

// ------------------------------------------------------------ fused kernel --
template <int FMT, bool FMA>
__global__ void __launch_bounds__(HT, 4)
hash_fused_kernel(const uint8_t *__restrict__ frames, size_t n, uint8_t *__restrict__ hashes,
                  float *__restrict__ quality, float *__restrict__ thresholds,
                  float *__restrict__ pdqf_out)
{
    using P = Px<FMT, FMA>;
    constexpr size_t ROW = 512 * P::BPP;   // bytes per input row
    __shared__ HashSmem sm;
    const int tid = threadIdx.x;
    const int j = tid < 127 ? tid : 126;  // thread 127 shadows column 126 (never stored)
    load_dt(sm, tid);
    __syncthreads();

    for (size_t f = blockIdx.x; f < n; f += gridDim.x) {
        const uint8_t *frame = frames + f * (512 * ROW);
        // window of output (oi, j): rows 4oi+2 .. 4oi+6, columns 4j+2 .. 4j+6
        const uint8_t *col = frame + 4 * j * P::BPP;
        if (tid == 0) {
#pragma unroll 1
            for (int a = 0; a < PREFETCH_AHEAD; a++)
                prefetch_l2(frame + (size_t)(4 * a + 2) * ROW, 4 * ROW);
        }
        float acc[16][4];
#pragma unroll
        for (int k = 0; k < 16; k++)
#pragma unroll
            for (int r = 0; r < 4; r++)
                acc[k][r] = 0.0f;

        // first window row (input row 2), carried between iterations afterwards
        Win5 carry = P::load(col + 2 * ROW);

        auto one_row = [&](int oi, auto rtag) {
            constexpr int R = decltype(rtag)::value;  // oi & 3
            const uint8_t *base = col + (size_t)(4 * oi + 3) * ROW;
            const Win5 w1 = P::load(base);
            const Win5 w2 = P::load(base + ROW);
            const Win5 w3 = P::load(base + 2 * ROW);
            const Win5 w4 = P::load(base + 3 * ROW);
            if (tid == 0 && oi + PREFETCH_AHEAD < 127)
                prefetch_l2(frame + (size_t)(4 * (oi + PREFETCH_AHEAD) + 3) * ROW, 4 * ROW);
            float sum = 0.0f;
            sum = tent_row<1>(sum, carry);
            sum = tent_row<2>(sum, w1);
            sum = tent_row<3>(sum, w2);
            sum = tent_row<4>(sum, w3);
            sum = tent_row<5>(sum, w4);
            carry = w4;
            rank1<R>(acc, sm.dt + oi * 16, sum);
        };
        using T0 = std::integral_constant<int, 0>;
        using T1 = std::integral_constant<int, 1>;
        using T2 = std::integral_constant<int, 2>;
        using T3 = std::integral_constant<int, 3>;
#pragma unroll 1
        for (int oi = 0; oi < 124; oi += 4) {
            one_row(oi, T0{});
            one_row(oi + 1, T1{});
            one_row(oi + 2, T2{});
            one_row(oi + 3, T3{});
        }
        one_row(124, T0{});
        one_row(125, T1{});
        one_row(126, T2{});

        __syncthreads();  // previous frame's epilogue has finished with sm.t / sm.p
        if (tid < 127) {
#pragma unroll
            for (int k = 0; k < 16; k++)
                sm.t[k * TSTRIDE + tid] =
                    __fadd_rn(__fadd_rn(__fadd_rn(acc[k][0], acc[k][1]), acc[k][2]), acc[k][3]);
        }
        __syncthreads();
        dct_stage2(sm, tid);
        __syncthreads();
        if (pdqf_out) {
            pdqf_out[f * 256 + tid] = sm.p[tid];
            pdqf_out[f * 256 + 128 + tid] = sm.p[128 + tid];
        }
        dihedral_epilogue(sm, tid, hashes + f * 256, quality ? quality + f * 8 : nullptr,
                          thresholds ? thresholds + f * 8 : nullptr, false, 0.0f);
    }
}

// ------------------------------------------------- per-stage (Kernel trait) --
// Kernel::jarosz_compress: one thread per output pixel, same arithmetic as above.
__global__ void __launch_bounds__(128)
jarosz_kernel(const float *__restrict__ frames, float *__restrict__ buf1)
{
    const size_t f = blockIdx.y;
    const int oi = blockIdx.x, oj = threadIdx.x;  // grid.x = 128 rows, 128 threads
    float *out = buf1 + f * (128 * 128) + oi * 128 + oj;
    if (oi >= 127 || oj >= 127) {
        *out = 0.0f;  // zero padding of the 128x128 buffer (x86.rs:756-757, lib.rs:296-297)
        return;
    }
    using P = Px<YUME_PDQ_CUDA_PIXEL_F32, false>;
    const uint8_t *base = reinterpret_cast<const uint8_t *>(frames + f * (512 * 512) + (size_t)(4 * oi + 2) * 512 + 4 * oj);
    float sum = 0.0f;
    sum = tent_row<1>(sum, P::load(base));
    sum = tent_row<2>(sum, P::load(base + 2048));
    sum = tent_row<3>(sum, P::load(base + 2 * 2048));
    sum = tent_row<4>(sum, P::load(base + 3 * 2048));
    sum = tent_row<5>(sum, P::load(base + 4 * 2048));
    *out = sum;
}

// Kernel::dct2d on a 128x128 (127x127 valid) buffer
__global__ void __launch_bounds__(HT)
dct2d_kernel(const float *__restrict__ buf1, float *__restrict__ pdqf)
{
    __shared__ HashSmem sm;
    const int tid = threadIdx.x;
    const int j = tid < 127 ? tid : 126;
    load_dt(sm, tid);
    __syncthreads();
    const float *b = buf1 + (size_t)blockIdx.x * (128 * 128) + j;
    float acc[16][4];
#pragma unroll
    for (int k = 0; k < 16; k++)
#pragma unroll
        for (int r = 0; r < 4; r++)
            acc[k][r] = 0.0f;
#pragma unroll 1
    for (int k2 = 0; k2 < 124; k2 += 4) {
        rank1<0>(acc, sm.dt + (k2 + 0) * 16, __ldg(b + (k2 + 0) * 128));
        rank1<1>(acc, sm.dt + (k2 + 1) * 16, __ldg(b + (k2 + 1) * 128));
        rank1<2>(acc, sm.dt + (k2 + 2) * 16, __ldg(b + (k2 + 2) * 128));
        rank1<3>(acc, sm.dt + (k2 + 3) * 16, __ldg(b + (k2 + 3) * 128));
    }
    rank1<0>(acc, sm.dt + 124 * 16, __ldg(b + 124 * 128));
    rank1<1>(acc, sm.dt + 125 * 16, __ldg(b + 125 * 128));
    rank1<2>(acc, sm.dt + 126 * 16, __ldg(b + 126 * 128));
    if (tid < 127) {
#pragma unroll
        for (int k = 0; k < 16; k++)
            sm.t[k * TSTRIDE + tid] =
                __fadd_rn(__fadd_rn(__fadd_rn(acc[k][0], acc[k][1]), acc[k][2]), acc[k][3]);
    }
    __syncthreads();
    dct_stage2(sm, tid);
    __syncthreads();
    pdqf[(size_t)blockIdx.x * 256 + tid] = sm.p[tid];
    pdqf[(size_t)blockIdx.x * 256 + 128 + tid] = sm.p[128 + tid];
}

// what: 0 gradients(+quality), 1 quantize, 2 threshold_2d_f32, 3 dihedrals
__global__ void __launch_bounds__(HT)
pdqf_ops_kernel(const float *__restrict__ pdqf, int what, const float *__restrict__ thr_in,
                float *__restrict__ f_out0, float *__restrict__ f_out1,
                uint8_t *__restrict__ hashes)
{
    __shared__ HashSmem sm;
    const int tid = threadIdx.x;
    const size_t f = blockIdx.x;
    sm.p[tid] = pdqf[f * 256 + tid];
    sm.p[128 + tid] = pdqf[f * 256 + 128 + tid];
    __syncthreads();
    if (what == 0) {
        if (tid == 0) {
            float g = gradient_sum(sm.p, 0);
            if (f_out0)
                f_out0[f] = g;
            if (f_out1)
                f_out1[f] = adjust_quality(g);
        }
    } else if (what == 1) {
        if (tid < 32) {
            float med = warp_lower_median(sm.p, 0, tid);
            if (tid == 0) {
                sm.thr[0] = med;
                f_out0[f] = med;
            }
        }
        __syncthreads();
        write_hash_bits(sm.p, 0, sm.thr[0], hashes + f * 32, tid);
    } else if (what == 2) {
        write_hash_bits(sm.p, 0, thr_in[f], hashes + f * 32, tid);
    } else {
        dihedral_epilogue(sm, tid, hashes + f * 256, f_out0 ? f_out0 + f * 8 : nullptr,
                          f_out1 ? f_out1 + f * 8 : nullptr, true, thr_in[f]);
    }
}

__global__ void __launch_bounds__(256)
pdqf_transform_kernel(float *__restrict__ pdqf, int op)
{
    __shared__ float p[256];
    const int tid = threadIdx.x, i = tid >> 4, j = tid & 15;
    float *base = pdqf + (size_t)blockIdx.x * 256;
    p[tid] = base[tid];
    __syncthreads();
    float v = p[tid];
    switch (op) {
    case YUME_PDQ_CUDA_OP_T: v = p[j * 16 + i]; break;
    case YUME_PDQ_CUDA_OP_NEGATE_ALT_COLS_FALSE: if (!(j & 1)) v = -v; break;
    case YUME_PDQ_CUDA_OP_NEGATE_ALT_COLS_TRUE: if (j & 1) v = -v; break;
    case YUME_PDQ_CUDA_OP_NEGATE_ALT_ROWS_TRUE: if (!(i & 1)) v = -v; break;
    case YUME_PDQ_CUDA_OP_NEGATE_ALT_ROWS_FALSE: if (i & 1) v = -v; break;
    case YUME_PDQ_CUDA_OP_NEGATE_OFF_DIAGONALS: if ((j - i) & 1) v = -v; break;
    default: break;
    }
    base[tid] = v;
}

// one pixel to f32 luma with explicit coefficients (the trait's const generics R_COEFF/G_COEFF/B_COEFF)
template <bool FMA>
__device__ __forceinline__ float luma_c(float r, float g, float b, float cr, float cg, float cb)
{
    if (FMA)
        return __fmaf_rn(b, cb, __fmaf_rn(g, cg, __fmul_rn(r, cr)));
    return __fadd_rn(__fadd_rn(__fmul_rn(cr, r), __fmul_rn(cg, g)), __fmul_rn(cb, b));
}

template <int FMT, bool FMA>
__global__ void __launch_bounds__(256)
cvt_luma8f_kernel(const uint8_t *__restrict__ px, size_t n_pixels, float cr, float cg, float cb,
                  float *__restrict__ out)
{
    constexpr int BPP = FMT == YUME_PDQ_CUDA_PIXEL_LUMA8 ? 1 : FMT == YUME_PDQ_CUDA_PIXEL_RGB8 ? 3 : 4;
    const size_t groups = (n_pixels + 3) / 4;
    const bool aligned = (reinterpret_cast<uintptr_t>(px) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups;
         g += (size_t)gridDim.x * blockDim.x) {
        const size_t p0 = 4 * g;
        if (aligned && p0 + 4 <= n_pixels) {
            float4 v;
            if (FMT == YUME_PDQ_CUDA_PIXEL_LUMA8) {
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(px) + g);
                v = make_float4(u8f(w, 0), u8f(w, 1), u8f(w, 2), u8f(w, 3));
            } else if (FMT == YUME_PDQ_CUDA_PIXEL_RGB8) {
                const uint32_t *q = reinterpret_cast<const uint32_t *>(px) + 3 * g;
                const uint32_t a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
                v = make_float4(luma_c<FMA>(u8f(a, 0), u8f(a, 1), u8f(a, 2), cr, cg, cb),
                                luma_c<FMA>(u8f(a, 3), u8f(b, 0), u8f(b, 1), cr, cg, cb),
                                luma_c<FMA>(u8f(b, 2), u8f(b, 3), u8f(c, 0), cr, cg, cb),
                                luma_c<FMA>(u8f(c, 1), u8f(c, 2), u8f(c, 3), cr, cg, cb));
            } else {
                const uint4 w = __ldg(reinterpret_cast<const uint4 *>(px) + g);
                v = make_float4(luma_c<FMA>(u8f(w.x, 0), u8f(w.x, 1), u8f(w.x, 2), cr, cg, cb),
                                luma_c<FMA>(u8f(w.y, 0), u8f(w.y, 1), u8f(w.y, 2), cr, cg, cb),
                                luma_c<FMA>(u8f(w.z, 0), u8f(w.z, 1), u8f(w.z, 2), cr, cg, cb),
                                luma_c<FMA>(u8f(w.w, 0), u8f(w.w, 1), u8f(w.w, 2), cr, cg, cb));
            }
            reinterpret_cast<float4 *>(out)[g] = v;
        } else {
            for (size_t p = p0; p < n_pixels && p < p0 + 4; p++) {
                const uint8_t *b = px + p * BPP;
                out[p] = FMT == YUME_PDQ_CUDA_PIXEL_LUMA8
                             ? (float)b[0]
                             : luma_c<FMA>((float)b[0], (float)b[1], (float)b[2], cr, cg, cb);
            }
        }
    }
}

int upload_luts()
{
    // once per device and process (the symbols live in the device's context, not in the calling thread)
    static std::atomic<int> done[64];
    static std::mutex mu;
    int dev = 0;
    YPDQ_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64)
        return YUME_PDQ_CUDA_E_INVALID;
    if (done[dev].load(std::memory_order_acquire))
        return 0;
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev].load(std::memory_order_acquire))
        return 0;
    static uint32_t dt[127 * 16];
    for (int k = 0; k < 16; k++)
        for (int k2 = 0; k2 < 127; k2++)
            dt[k2 * 16 + k] = YUME_PDQ_DCT_MATRIX_BITS[k * 127 + k2];
    YPDQ_CUDA_TRY(cudaMemcpyToSymbol(g_dt, dt, sizeof dt));
    YPDQ_CUDA_TRY(cudaMemcpyToSymbol(c_tent_bits, YUME_PDQ_TENT_WEIGHT_BITS, sizeof(uint32_t) * 49));
    done[dev].store(1, std::memory_order_release);
    return 0;
}

// The `threshold` value the reference's AVX2 / AVX-512 kernels report (x86.rs:440-621): not the lower median but the
// point where a mean-of-the-residual-side bisection stops, anywhere in [sorted[127], sorted[128]).  The hash bits are
// the same; callers that print the float see a different number (README.md:164-168: 12.554 where Torben gives
// 11.907).  Lanes 0..7 of one warp play the eight lanes of a __m256, in the reference's operation order.
// one full warp; the result is valid in lane 0
__device__ float avx2_style_threshold(const float *pdqf)
{
    const unsigned full = 0xffffffffu;
    const int l = threadIdx.x & 7;
    auto hsum8 = [&](float a) {     // horizontal!(sum ..): ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7)) in lane 0
        a = __fadd_rn(a, __shfl_sync(full, a, (l + 1) & 7, 8));
        a = __fadd_rn(a, __shfl_sync(full, a, (l + 2) & 7, 8));
        a = __fadd_rn(a, __shfl_sync(full, a, (l + 4) & 7, 8));
        return __shfl_sync(full, a, 0, 8);
    };
    float mx = -3.40282347e+38f, mn = 3.40282347e+38f;
    for (int c = 0; c < 32; c++) {
        const float v = pdqf[8 * c + l];
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    for (int d = 1; d < 8; d <<= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(full, mx, d, 8));
        mn = fminf(mn, __shfl_xor_sync(full, mn, d, 8));
    }
    float guess = __fmul_rn(__fadd_rn(mn, mx), 0.5f);
    uint32_t num_over = 0, num_under = 0;
    for (int iter = 0; iter < 32; iter++) {
        float rs_gt = 0.0f, rs_lt = 0.0f;
        uint32_t gt = 0, lt = 0;
        for (int c = 0; c < 32; c++) {
            const float v = pdqf[8 * c + l];
            const bool mgt = v > guess && v <= mx, mlt = v < guess && v >= mn;
            gt += mgt;
            lt += mlt;
            rs_gt = __fadd_rn(rs_gt, mgt ? v : 0.0f);
            rs_lt = __fadd_rn(rs_lt, mlt ? v : 0.0f);
        }
        for (int d = 1; d < 8; d <<= 1) {
            gt += __shfl_xor_sync(full, gt, d, 8);
            lt += __shfl_xor_sync(full, lt, d, 8);
        }
        const float s_gt = hsum8(rs_gt), s_lt = hsum8(rs_lt);
        if (gt + num_over > 128u) {
            num_under += lt;
            mn = guess;
            guess = __fdiv_rn(s_gt, (float)gt);
        } else if (lt + num_under > 128u) {
            num_over += gt;
            mx = guess;
            guess = __fdiv_rn(s_lt, (float)lt);
        } else {
            break;
        }
    }
    return guess;
}

__global__ void __launch_bounds__(32)
avx2_style_threshold_kernel(const float *__restrict__ pdqf, float *__restrict__ out)
{
    const float guess = avx2_style_threshold(pdqf);
    if (threadIdx.x == 0)
        *out = guess;
}

// ---------------------------------------------------------------- one frame, latency path --
// hash_get_threshold (src/lib.rs:427-455) for ONE frame in ONE launch, built for the latency of the reference's
// single-frame C ABI (yume_pdq_hash_smart_kernel) rather than for throughput:
//   * 128 CTAs, one per row of the 128 x 128 tent buffer (the fused batch kernel streams a frame through one CTA);
//   * the last CTA to finish (device-scope counter) stages the tent buffer in shared memory and runs dct2d, the lower
//     median, the gradient sum and the hash bits exactly as the per-stage kernels do;
//   * results go straight to mapped host memory, then the sequence number: the host polls it, no D2H copy, no stream
//     synchronisation.
// (Reading the frame in place from pinned host memory was measured too: 1 MiB of 16-byte loads over PCIe takes ~40 us,
// 26 GB/s, where the copy engine needs 19 us; and letting the kernel chase the host's copy chunk by chunk slowed the
// host's copy down to 150 us -- the frame is therefore always in device memory here.)
// SM clock for the diagnostic trace; the memory clobber keeps the read where it is written (between barriers)
__device__ __forceinline__ long long one_clock()
{
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
}

constexpr int ONE_T = 512;      // threads per CTA: the tent row uses 128 of them, the finishing CTA all

__global__ void __launch_bounds__(ONE_T)
hash_one_kernel(const float *__restrict__ frame, float *buf1_dev, unsigned int *done, OneFrameOut *out, uint32_t seq,
                int avx2_style)
{
    extern __shared__ __align__(16) float s_b[];      // finishing CTA: the 128 x 128 tent buffer, later the gradient terms
    __shared__ HashSmem sm;
    __shared__ unsigned int s_last;
    const int tid = threadIdx.x, oi = blockIdx.x, lane = tid & 31;
    const long long t_start = one_clock();
    if (tid < 128) {
        float sum = 0.0f;      // row 127 and column 127: zero padding (x86.rs:756-757, lib.rs:296-297)
        if (oi < 127) {
            // columns 4j .. 4j+3 from this thread's load, 4j+4 .. 4j+6 from the right neighbour's
            const float *row = frame + (size_t)(4 * oi + 2) * 512 + 4 * tid;
            float4 a[5], b[5];
#pragma unroll
            for (int r = 0; r < 5; r++)
                a[r] = __ldg(reinterpret_cast<const float4 *>(row + r * 512));
#pragma unroll
            for (int r = 0; r < 5; r++) {
                b[r].x = __shfl_down_sync(0xffffffffu, a[r].x, 1);
                b[r].y = __shfl_down_sync(0xffffffffu, a[r].y, 1);
                b[r].z = __shfl_down_sync(0xffffffffu, a[r].z, 1);
            }
            if (lane == 31 && tid < 127) {
#pragma unroll
                for (int r = 0; r < 5; r++)
                    b[r] = __ldg(reinterpret_cast<const float4 *>(row + r * 512 + 4));
            }
            if (tid < 127) {
                sum = tent_row<1>(sum, Win5{a[0].z, a[0].w, b[0].x, b[0].y, b[0].z});
                sum = tent_row<2>(sum, Win5{a[1].z, a[1].w, b[1].x, b[1].y, b[1].z});
                sum = tent_row<3>(sum, Win5{a[2].z, a[2].w, b[2].x, b[2].y, b[2].z});
                sum = tent_row<4>(sum, Win5{a[3].z, a[3].w, b[3].x, b[3].y, b[3].z});
                sum = tent_row<5>(sum, Win5{a[4].z, a[4].w, b[4].x, b[4].y, b[4].z});
            }
        }
        buf1_dev[oi * 128 + tid] = sum;
        out->buf1[oi * 128 + tid] = sum;
        // system scope: this row of buf1 sits in HOST memory and must be there before the finishing CTA's sequence number
        __threadfence_system();
    }
    __syncthreads();
    if (tid == 0)
        s_last = atomicAdd(done, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (!s_last)
        return;
    if (tid == 0) {
        *done = 0;
        out->trace[0] = t_start;
        out->trace[1] = one_clock();
    }
    __threadfence();
    {
        const float4 *g4 = reinterpret_cast<const float4 *>(buf1_dev);
        float4 *s4 = reinterpret_cast<float4 *>(s_b);
        float4 v[8];      // 128 x 128 / 4 / 512 = 8 loads per thread, all in flight
#pragma unroll
        for (int i = 0; i < 8; i++)
            v[i] = __ldcg(g4 + tid + i * ONE_T);
#pragma unroll
        for (int i = 0; i < 8; i++)
            s4[tid + i * ONE_T] = v[i];
    }
    {
        const float4 *src = reinterpret_cast<const float4 *>(g_dt);
        float4 *dst = reinterpret_cast<float4 *>(sm.dt);
        if (tid < 127 * 4)
            dst[tid] = src[tid];
    }
    __syncthreads();
    if (tid == 0)
        out->trace[2] = one_clock();
    {
        // dct2d stage 1 as in dct2d_kernel, the 16 DCT rows split over four thread groups: thread (j, kq) carries rows
        // 4 kq .. 4 kq + 3 of column j -- every (k, j, partial sum) chain adds the same terms in the same order
        const int j = (tid & 127) < 127 ? (tid & 127) : 126, kq = tid >> 7;
        const float *b = s_b + j;
        const float *dq = sm.dt + 4 * kq;
        float acc[4][4];
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int r = 0; r < 4; r++)
                acc[k][r] = 0.0f;
        auto upd = [&](int k2, int r, float bv) {
            const float4 d = *reinterpret_cast<const float4 *>(dq + k2 * 16);
            acc[0][r] = __fadd_rn(acc[0][r], __fmul_rn(d.x, bv));
            acc[1][r] = __fadd_rn(acc[1][r], __fmul_rn(d.y, bv));
            acc[2][r] = __fadd_rn(acc[2][r], __fmul_rn(d.z, bv));
            acc[3][r] = __fadd_rn(acc[3][r], __fmul_rn(d.w, bv));
        };
#pragma unroll 4
        for (int k2 = 0; k2 < 124; k2 += 4) {
            upd(k2 + 0, 0, b[(k2 + 0) * 128]);
            upd(k2 + 1, 1, b[(k2 + 1) * 128]);
            upd(k2 + 2, 2, b[(k2 + 2) * 128]);
            upd(k2 + 3, 3, b[(k2 + 3) * 128]);
        }
        upd(124, 0, b[124 * 128]);
        upd(125, 1, b[125 * 128]);
        upd(126, 2, b[126 * 128]);
        if ((tid & 127) < 127) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                sm.t[(4 * kq + k) * TSTRIDE + j] = __fadd_rn(__fadd_rn(__fadd_rn(acc[k][0], acc[k][1]), acc[k][2]), acc[k][3]);
        }
    }
    __syncthreads();
    if (tid == 0)
        out->trace[3] = one_clock();
    if (tid < HT)
        dct_stage2(sm, tid);
    __syncthreads();
    if (tid == 0)
        out->trace[4] = one_clock();
    if (tid < 256)
        out->pdqf[tid] = sm.p[tid];
    // sum_of_gradients (mod.rs:617-642): the 480 terms in parallel, then ONE thread adds them in the reference's order
    if (tid < 240)
        s_b[tid] = fabsf(__fsub_rn(sm.p[tid], sm.p[tid + 16]));                              // (i, j) - (i + 1, j), i < 15
    else if (tid < 480) {
        const int t = tid - 240, i = t / 15, j = t - 15 * i;
        s_b[tid] = fabsf(__fsub_rn(sm.p[i * 16 + j], sm.p[i * 16 + j + 1]));                 // (i, j) - (i, j + 1), j < 15
    }
    __syncthreads();
    if (tid < 32) {
        const float med = warp_lower_median(sm.p, 0, tid);
        if (tid == 0) {
            sm.thr[0] = med;
            out->thr = med;
            out->trace[5] = one_clock();
        }
    } else if (tid == 32) {
        float g = 0.0f;
#pragma unroll 16
        for (int t = 0; t < 480; t++)
            g = __fadd_rn(g, s_b[t]);
        out->gradient = g;
        out->quality = adjust_quality(g);
        long long t;
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "f"(g) : "memory");
        out->trace[7] = t;
    } else if (tid >= 64 && tid < 96 && avx2_style) {
        const float t = avx2_style_threshold(sm.p);
        if (tid == 64)
            out->thr_avx2 = t;
    }
    __syncthreads();
    if (tid < HT)
        write_hash_bits(sm.p, 0, sm.thr[0], out->hash, tid);
    if (tid == 0)
        out->trace[6] = one_clock();
    __threadfence_system();
    __syncthreads();
    if (tid == 0)
        st_release_sys(&out->seq, seq);
}

// visit_dihedrals (lib.rs:358-420) for ONE PDQF as a latency path: the PDQF arrives in the launch parameters, the 7
// transformed PDQFs, 8 hashes and 8 qualities go straight to mapped host memory, then the sequence number
__global__ void __launch_bounds__(HT)
dihedrals_one_kernel(const __grid_constant__ VisitIn in, float thr0, VisitOut *out, uint32_t seq)
{
    __shared__ HashSmem sm;
    const int tid = threadIdx.x;
    sm.p[tid] = in.p[tid];
    sm.p[128 + tid] = in.p[128 + tid];
    __syncthreads();
#pragma unroll 1
    for (int d = 1; d < 8; d++) {
        out->states[(d - 1) * 256 + tid] = dihedral_value(sm.p, d, tid >> 4, tid & 15);
        out->states[(d - 1) * 256 + 128 + tid] = dihedral_value(sm.p, d, (128 + tid) >> 4, tid & 15);
    }
    dihedral_epilogue(sm, tid, out->hashes, out->quality, nullptr, true, thr0);
    __threadfence_system();
    __syncthreads();
    if (tid == 0)
        st_release_sys(&out->seq, seq);
}

}  // namespace

int launch_dihedrals_one(const VisitIn &in, float thr0, VisitOut *out, uint32_t seq, cudaStream_t s)
{
    dihedrals_one_kernel<<<1, HT, 0, s>>>(in, thr0, out, seq);
    YPDQ_LAUNCH_CHECK();
    return 0;
}

int launch_hash_batch(const void *frames, int pixel_format, size_t n, uint8_t *hashes, float *quality,
                      float *thresholds, float *pdqf, cudaStream_t s)
{
    if (n == 0)
        return 0;
    int rc = upload_luts();
    if (rc)
        return rc;
    const unsigned grid = (unsigned)std::min<size_t>(n, (size_t)1 << 30);
    const uint8_t *fr = static_cast<const uint8_t *>(frames);
#define YPDQ_HASH_CASE(FMT, FMA)                                                                       \
    case (FMT) | ((FMA) ? YUME_PDQ_CUDA_PIXEL_FMA : 0):                                                \
        hash_fused_kernel<FMT, FMA><<<grid, HT, 0, s>>>(fr, n, hashes, quality, thresholds, pdqf);     \
        break
    switch (pixel_format) {
        YPDQ_HASH_CASE(YUME_PDQ_CUDA_PIXEL_F32, false);
        YPDQ_HASH_CASE(YUME_PDQ_CUDA_PIXEL_LUMA8, false);
        YPDQ_HASH_CASE(YUME_PDQ_CUDA_PIXEL_RGB8, false);
        YPDQ_HASH_CASE(YUME_PDQ_CUDA_PIXEL_RGB8, true);
        YPDQ_HASH_CASE(YUME_PDQ_CUDA_PIXEL_RGBA8, false);
        YPDQ_HASH_CASE(YUME_PDQ_CUDA_PIXEL_RGBA8, true);
    default:
        return YUME_PDQ_CUDA_E_INVALID;
    }
#undef YPDQ_HASH_CASE
    YPDQ_LAUNCH_CHECK();
    return 0;
}

// Kernel::cvt_rgb8_to_luma8f / cvt_rgba8_to_luma8f (mod.rs:303-340) and the LUMA8 `as f32` of main.rs:508-510
// over n_pixels pixels (any number of rows); four pixels per thread, byte-wise tail.
int launch_cvt_luma8f(const void *pixels, int pixel_format, size_t n_pixels, const float *coeff, float *out,
                      cudaStream_t s)
{
    if (n_pixels == 0)
        return 0;
    const int fmt = pixel_format & ~YUME_PDQ_CUDA_PIXEL_FMA;
    const bool fma = (pixel_format & YUME_PDQ_CUDA_PIXEL_FMA) != 0;
    if (fmt != YUME_PDQ_CUDA_PIXEL_LUMA8 && fmt != YUME_PDQ_CUDA_PIXEL_RGB8 && fmt != YUME_PDQ_CUDA_PIXEL_RGBA8)
        return YUME_PDQ_CUDA_E_INVALID;
    const float cr = coeff ? coeff[0] : 0.299f, cg = coeff ? coeff[1] : 0.587f, cb = coeff ? coeff[2] : 0.114f;
    const size_t groups = (n_pixels + 3) / 4;
    const unsigned grid = (unsigned)std::min<size_t>((groups + 255) / 256, 148 * 16);
    const uint8_t *px = static_cast<const uint8_t *>(pixels);
    if (fmt == YUME_PDQ_CUDA_PIXEL_LUMA8)
        cvt_luma8f_kernel<YUME_PDQ_CUDA_PIXEL_LUMA8, false><<<grid, 256, 0, s>>>(px, n_pixels, cr, cg, cb, out);
    else if (fmt == YUME_PDQ_CUDA_PIXEL_RGB8 && !fma)
        cvt_luma8f_kernel<YUME_PDQ_CUDA_PIXEL_RGB8, false><<<grid, 256, 0, s>>>(px, n_pixels, cr, cg, cb, out);
    else if (fmt == YUME_PDQ_CUDA_PIXEL_RGB8)
        cvt_luma8f_kernel<YUME_PDQ_CUDA_PIXEL_RGB8, true><<<grid, 256, 0, s>>>(px, n_pixels, cr, cg, cb, out);
    else if (!fma)
        cvt_luma8f_kernel<YUME_PDQ_CUDA_PIXEL_RGBA8, false><<<grid, 256, 0, s>>>(px, n_pixels, cr, cg, cb, out);
    else
        cvt_luma8f_kernel<YUME_PDQ_CUDA_PIXEL_RGBA8, true><<<grid, 256, 0, s>>>(px, n_pixels, cr, cg, cb, out);
    YPDQ_LAUNCH_CHECK();
    return 0;
}

int launch_avx2_style_threshold(const float *pdqf, float *out, cudaStream_t s)
{
    avx2_style_threshold_kernel<<<1, 32, 0, s>>>(pdqf, out);
    YPDQ_LAUNCH_CHECK();
    return 0;
}

int launch_hash_one(const float *frame, float *buf1_dev, unsigned int *done, OneFrameOut *out, uint32_t seq, bool avx2_style,
                    cudaStream_t s)
{
    int rc = upload_luts();
    if (rc)
        return rc;
    constexpr int dyn = 128 * 128 * 4;
    static std::atomic<int> attr_set[64];
    int dev = 0;
    YPDQ_CUDA_TRY(cudaGetDevice(&dev));
    if (!attr_set[dev & 63].load(std::memory_order_acquire)) {
        YPDQ_CUDA_TRY(cudaFuncSetAttribute(hash_one_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
        attr_set[dev & 63].store(1, std::memory_order_release);
    }
    hash_one_kernel<<<128, ONE_T, dyn, s>>>(frame, buf1_dev, done, out, seq, avx2_style ? 1 : 0);
    YPDQ_LAUNCH_CHECK();
    return 0;
}

int launch_jarosz(const float *frames, size_t n, float *buf1, cudaStream_t s)
{
    int rc = upload_luts();
    if (rc)
        return rc;
    for (size_t f0 = 0; f0 < n; f0 += 65535) {
        unsigned cnt = (unsigned)std::min<size_t>(65535, n - f0);
        jarosz_kernel<<<dim3(128, cnt), 128, 0, s>>>(frames + f0 * 512 * 512, buf1 + f0 * 128 * 128);
        YPDQ_LAUNCH_CHECK();
    }
    return 0;
}

int launch_dct2d(const float *buf1, size_t n, float *pdqf, cudaStream_t s)
{
    if (n == 0)
        return 0;
    int rc = upload_luts();
    if (rc)
        return rc;
    dct2d_kernel<<<(unsigned)n, HT, 0, s>>>(buf1, pdqf);
    YPDQ_LAUNCH_CHECK();
    return 0;
}

int launch_gradients(const float *pdqf, size_t n, float *gradient, float *quality, cudaStream_t s)
{
    if (n == 0)
        return 0;
    pdqf_ops_kernel<<<(unsigned)n, HT, 0, s>>>(pdqf, 0, nullptr, gradient, quality, nullptr);
    YPDQ_LAUNCH_CHECK();
    return 0;
}

int launch_quantize(const float *pdqf, size_t n, float *thresholds, uint8_t *hashes, cudaStream_t s)
{
    if (n == 0)
        return 0;
    pdqf_ops_kernel<<<(unsigned)n, HT, 0, s>>>(pdqf, 1, nullptr, thresholds, nullptr, hashes);
    YPDQ_LAUNCH_CHECK();
    return 0;
}

int launch_threshold(const float *pdqf, size_t n, const float *thresholds, uint8_t *hashes,
                     cudaStream_t s)
{
    if (n == 0)
        return 0;
    pdqf_ops_kernel<<<(unsigned)n, HT, 0, s>>>(pdqf, 2, thresholds, nullptr, nullptr, hashes);
    YPDQ_LAUNCH_CHECK();
    return 0;
}

int launch_transform(float *pdqf, size_t n, int op, cudaStream_t s)
{
    if (n == 0)
        return 0;
    pdqf_transform_kernel<<<(unsigned)n, 256, 0, s>>>(pdqf, op);
    YPDQ_LAUNCH_CHECK();
    return 0;
}

int launch_dihedrals(const float *pdqf, const float *thr0, size_t n, uint8_t *hashes, float *quality,
                     float *thresholds, cudaStream_t s)
{
    if (n == 0)
        return 0;
    pdqf_ops_kernel<<<(unsigned)n, HT, 0, s>>>(pdqf, 3, thr0, quality, thresholds, hashes);
    YPDQ_LAUNCH_CHECK();
    return 0;
}

}  // namespace ypdq
