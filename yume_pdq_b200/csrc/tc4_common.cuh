// tc4_common.cuh -- mbarrier / bulk-copy / tcgen05 helpers shared by the packed-FP4 tensor matchers
// (matcher_tc4.cu, matcher_tc4x2.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace ypdq {
namespace tc4 {

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
__device__ __forceinline__ bool bar_try(uint64_t *b, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(s32(b)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void bar_wait(uint64_t *b, uint32_t parity)
{
    while (!bar_try(b, parity)) {
    }
}
// hot-loop forms on precomputed shared-window addresses (the generic->shared conversion costs an S2R + LEA
// per use otherwise); the wait is one PTX block, try_wait parks the warp in hardware between polls
__device__ __forceinline__ void bar_wait_addr(uint32_t addr, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "T4_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra T4_DONE;\n\t"
        "bra T4_WAIT;\n"
        "T4_DONE:\n\t}" ::"r"(addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bar_arrive_addr(uint32_t addr)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
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

// 8 bits -> 8 e2m1 nibbles: bit i -> nibble i; bit 1 -> 0x2 (+1.0), bit 0 -> 0xA (-1.0)
__device__ __forceinline__ uint32_t expand8(uint32_t x)
{
    x = (x | (x << 12)) & 0x000F000Fu;
    x = (x | (x << 6)) & 0x03030303u;
    x = (x | (x << 3)) & 0x11111111u;
    return (x << 3) ^ 0xAAAAAAAAu;
}
// one 32-bit hash word -> one 16-byte K chunk (32 e2m1 elements)
__device__ __forceinline__ uint4 expand32(uint32_t w)
{
    return make_uint4(expand8(w & 255u), expand8((w >> 8) & 255u), expand8((w >> 16) & 255u), expand8(w >> 24));
}

struct Tc4Sink {
    YumePDQ_Match *out;
    uint32_t cap;
    uint32_t *count;
    unsigned long long *min_key;
    uint64_t row_offset;
    uint64_t q;
    int mode;
    uint32_t needle_base;   // added to reported needle indices (LIST mode): pipelined callers match in chunks
    const uint4 *needles_raw;   // q x 32 B: prefix-filter candidates are verified on the packed bits
    uint32_t threshold;
};

__device__ __forceinline__ void tc4_emit(const Tc4Sink &k, uint64_t local_row, uint32_t needle, uint32_t dist)
{
    uint32_t pos = atomicAdd(k.count, 1u);
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

// full 256-bit distance of a candidate, on the packed bits
__device__ __forceinline__ uint32_t hamming256(const uint4 &a0, const uint4 &a1, const uint4 &b0, const uint4 &b1)
{
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

#define T4_LD_X32(v, o, taddr)                                                                                 \
    asm volatile(                                                                                              \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                              \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                             \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"              \
        : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]),      \
          "=r"(v[o + 6]), "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]),    \
          "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15]), "=r"(v[o + 16]), "=r"(v[o + 17]), \
          "=r"(v[o + 18]), "=r"(v[o + 19]), "=r"(v[o + 20]), "=r"(v[o + 21]), "=r"(v[o + 22]), "=r"(v[o + 23]), \
          "=r"(v[o + 24]), "=r"(v[o + 25]), "=r"(v[o + 26]), "=r"(v[o + 27]), "=r"(v[o + 28]), "=r"(v[o + 29]), \
          "=r"(v[o + 30]), "=r"(v[o + 31])                                                                     \
        : "r"(taddr))
#define T4_LD_X64(v, o, taddr) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];" \
                 : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]), "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15]), "=r"(v[o + 16]), "=r"(v[o + 17]), "=r"(v[o + 18]), "=r"(v[o + 19]), "=r"(v[o + 20]), "=r"(v[o + 21]), "=r"(v[o + 22]), "=r"(v[o + 23]), "=r"(v[o + 24]), "=r"(v[o + 25]), "=r"(v[o + 26]), "=r"(v[o + 27]), "=r"(v[o + 28]), "=r"(v[o + 29]), "=r"(v[o + 30]), "=r"(v[o + 31]), "=r"(v[o + 32]), "=r"(v[o + 33]), "=r"(v[o + 34]), "=r"(v[o + 35]), "=r"(v[o + 36]), "=r"(v[o + 37]), "=r"(v[o + 38]), "=r"(v[o + 39]), "=r"(v[o + 40]), "=r"(v[o + 41]), "=r"(v[o + 42]), "=r"(v[o + 43]), "=r"(v[o + 44]), "=r"(v[o + 45]), "=r"(v[o + 46]), "=r"(v[o + 47]), "=r"(v[o + 48]), "=r"(v[o + 49]), "=r"(v[o + 50]), "=r"(v[o + 51]), "=r"(v[o + 52]), "=r"(v[o + 53]), "=r"(v[o + 54]), "=r"(v[o + 55]), "=r"(v[o + 56]), "=r"(v[o + 57]), "=r"(v[o + 58]), "=r"(v[o + 59]), "=r"(v[o + 60]), "=r"(v[o + 61]), "=r"(v[o + 62]), "=r"(v[o + 63]) \
                 : "r"(taddr))
#define T4_LD_X16(v, o, taddr)                                                                                 \
    asm volatile(                                                                                              \
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                              \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                       \
        : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]),      \
          "=r"(v[o + 6]), "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]),    \
          "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])                                   \
        : "r"(taddr))
#define T4_LD_X8(v, o, taddr)                                                                                  \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"               \
                 : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), \
                   "=r"(v[o + 6]), "=r"(v[o + 7])                                                              \
                 : "r"(taddr))
#define T4_LD_X4(v, o, taddr)                                                                                  \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"                               \
                 : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3])                              \
                 : "r"(taddr))

}  // namespace tc4
}  // namespace ypdq
