// tc4x2_test.cu -- bring-up harness (developer tool, not product) for the "two rows per accumulator" packing
// of the FP4 Hamming tile: one tcgen05.mma kind::mxf4 accumulator column holds
//     Z = 1.5 * 2^23  +  dot128(needle, rowA)  +  65536 * dot128(needle, rowB)
// built from 5 MMAs (K = 64 each): one constant "bias" MMA (A scale 2^22), two over rowA's 128-bit prefix (scale 1)
// and two over rowB's (A scale 2^16).  Every partial sum is an integer below 2^24, so f32 accumulation must be
// exact; the f32 bit pattern is then 0x4B400000 + dotA + 65536 * dotB, i.e. two s16 fields a SIMD max can test.
// The harness checks (1) exactness of the accumulation, (2) how many TMEM columns a scale-factor region needs
// (each region is filled with one byte value; the columns around it are poisoned with NaN scales), (3) that the A
// operand can be staged in TMEM (tcgen05.cp.128x256b, then tcgen05.mma with a TMEM address for A) with the same result.
// build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/tc4x2_test tools/tc4x2_test.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (uint64_t)((lbo >> 4) & 0x3FFF) << 16 | (uint64_t)((sbo >> 4) & 0x3FFF) << 32 |
           (uint64_t)1 << 46;
}

constexpr int M = 128, N = 240;
constexpr uint32_t LBO = 128;
constexpr int A_BYTES = M * 64, B_BYTES = N * 64, AC_BYTES = M * 32, BC_BYTES = N * 32;
constexpr uint32_t IDESC = (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((uint32_t)(M >> 4) << 24);

struct Params {
    uint32_t sf_b, sf_al, sf_ah, sf_ac;   // first TMEM column of each scale-factor region
    uint32_t width;                       // columns per region
    uint32_t bias_last;
    uint32_t a_tmem;                      // 1: every A operand is first copied to TMEM (tcgen05.cp) and read from there
};

__global__ void __launch_bounds__(128, 1)
tile_kernel(const uint8_t *__restrict__ tiles, uint32_t *__restrict__ d_out, Params p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem, *sB0 = sA + A_BYTES, *sB1 = sB0 + B_BYTES, *sAc = sB1 + B_BYTES, *sBc = sAc + AC_BYTES;
    constexpr int TOTAL = A_BYTES + 2 * B_BYTES + AC_BYTES + BC_BYTES;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + TOTAL);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < TOTAL / 16; i += 128)
        reinterpret_cast<uint4 *>(smem)[i] = reinterpret_cast<const uint4 *>(tiles)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    {
        // columns 448..511: poison (0xFF = NaN scale), then the four uniform regions
        const uint32_t base = tmem + (((uint32_t)warp * 32u) << 16);
        for (uint32_t c = 448; c < 512; c++) {
            uint32_t v = 0xFFFFFFFFu;
            if (c >= p.sf_b && c < p.sf_b + p.width) v = 0x7F7F7F7Fu;
            if (c >= p.sf_al && c < p.sf_al + p.width) v = 0x7F7F7F7Fu;
            if (c >= p.sf_ah && c < p.sf_ah + p.width) v = 0x8F8F8F8Fu;   // 2^16
            if (c >= p.sf_ac && c < p.sf_ac + p.width) v = 0x95959595u;   // 2^22
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(base + c), "r"(v) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (tid == 0) {
        uint32_t a_slot = 0;
        auto mma = [&](uint64_t da, uint64_t db, uint32_t sfa, uint32_t acc) {
            if (p.a_tmem) {
                // 128 rows x 256 bits (K = 64 e2m1) -> 8 TMEM columns; slots rotate over columns 448..479
                const uint32_t a_addr = tmem + 448 + 8 * (a_slot++ & 3u);
                asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(a_addr), "l"(da) : "memory");
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%5], [%6], p;\n\t}" ::"r"(tmem),
                    "r"(a_addr), "l"(db), "r"(IDESC), "r"(acc), "r"(tmem + sfa), "r"(tmem + p.sf_b)
                    : "memory");
                return;
            }
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(tmem),
                "l"(da), "l"(db), "r"(IDESC), "r"(acc), "r"(tmem + sfa), "r"(tmem + p.sf_b)
                : "memory");
        };
        uint32_t acc = 0;
        if (!p.bias_last) {
            mma(make_desc(s32(sAc), LBO, 256), make_desc(s32(sBc), LBO, 256), p.sf_ac, acc);
            acc = 1;
        }
        for (uint32_t k = 0; k < 2; k++, acc = 1)
            mma(make_desc(s32(sA) + k * 2 * LBO, LBO, 512), make_desc(s32(sB1) + k * 2 * LBO, LBO, 512), p.sf_ah, acc);
        for (uint32_t k = 0; k < 2; k++)
            mma(make_desc(s32(sA) + k * 2 * LBO, LBO, 512), make_desc(s32(sB0) + k * 2 * LBO, LBO, 512), p.sf_al, 1);
        if (p.bias_last)
            mma(make_desc(s32(sAc), LBO, 256), make_desc(s32(sBc), LBO, 256), p.sf_ac, 1);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
    }
    {
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(s32(bar)), "r"(0u) : "memory");
        } while (!done);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; j++)
            d_out[(size_t)tid * N + c0 + j] = v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// rows x 16 bytes of bits -> canonical K-major tile of +-1 e2m1 nibbles, 64 bytes per row
static void pack(const std::vector<uint8_t> &bits, int rows, uint8_t *tile)
{
    memset(tile, 0, (size_t)rows * 64);
    for (int r = 0; r < rows; r++)
        for (int k = 0; k < 128; k++) {
            const int bit = (bits[r * 16 + k / 8] >> (k % 8)) & 1;
            const size_t off = (size_t)(r / 8) * 512 + (size_t)(k / 32) * LBO + (r % 8) * 16 + (k % 32) / 2;
            tile[off] |= (bit ? 0x2 : 0xA) << (4 * (k & 1));
        }
}
// constant tile: element 0 of every row = nibble, the other 63 elements 0; 32 bytes per row
static void pack_const(int rows, uint8_t nibble, uint8_t *tile)
{
    memset(tile, 0, (size_t)rows * 32);
    for (int r = 0; r < rows; r++)
        tile[(size_t)(r / 8) * 256 + (r % 8) * 16] = nibble;
}

int main()
{
    srand(7);
    std::vector<uint8_t> a(M * 16), b0(N * 16), b1(N * 16);
    for (auto &x : a) x = rand();
    for (auto &x : b0) x = rand();
    for (auto &x : b1) x = rand();
    // extremes: exact copies and complements in either field, and both at once
    memcpy(&b0[3 * 16], &a[5 * 16], 16);                                   // (5,3): dotA = +128
    memcpy(&b1[4 * 16], &a[6 * 16], 16);                                   // (6,4): dotB = +128 (leaves the window)
    for (int i = 0; i < 16; i++) b1[9 * 16 + i] = ~a[7 * 16 + i];           // (7,9): dotB = -128 (leaves the window)
    for (int i = 0; i < 16; i++) b0[10 * 16 + i] = ~a[8 * 16 + i];          // (8,10): dotA = -128
    memcpy(&b0[11 * 16], &a[9 * 16], 16);
    memcpy(&b1[11 * 16], &a[9 * 16], 16);                                  // (9,11): both +128
    b1[12 * 16] = a[10 * 16] ^ 1;
    memcpy(&b1[12 * 16 + 1], &a[10 * 16 + 1], 15);                         // (10,12): dotB = 126
    constexpr int TOTAL = A_BYTES + 2 * B_BYTES + AC_BYTES + BC_BYTES;
    std::vector<uint8_t> tiles(TOTAL);
    pack(a, M, tiles.data());
    pack(b0, N, tiles.data() + A_BYTES);
    pack(b1, N, tiles.data() + A_BYTES + B_BYTES);
    pack_const(M, 0x3, tiles.data() + A_BYTES + 2 * B_BYTES);              // 1.5
    pack_const(N, 0x4, tiles.data() + A_BYTES + 2 * B_BYTES + AC_BYTES);   // 2.0
    uint8_t *dt;
    uint32_t *dd;
    CK(cudaMalloc(&dt, TOTAL));
    CK(cudaMalloc(&dd, M * N * 4));
    CK(cudaMemcpy(dt, tiles.data(), TOTAL, cudaMemcpyHostToDevice));
    const size_t smem = TOTAL + 64;
    CK(cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto dot = [&](const std::vector<uint8_t> &x, int m, const std::vector<uint8_t> &y, int n) {
        int d = 0;
        for (int w = 0; w < 16; w++) d += __builtin_popcount((uint8_t)(x[m * 16 + w] ^ y[n * 16 + w]));
        return 128 - 2 * d;
    };
    int rc = 0;
    const Params variants[] = {
        {480, 488, 496, 504, 8, 0, 0}, {480, 488, 496, 504, 8, 1, 0}, {480, 488, 496, 504, 8, 0, 1},
    };
    for (const Params &p : variants) {
        CK(cudaMemset(dd, 0, M * N * 4));
        tile_kernel<<<1, 128, smem>>>(dt, dd, p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return 2; }
        std::vector<uint32_t> got(M * N);
        CK(cudaMemcpy(got.data(), dd, M * N * 4, cudaMemcpyDeviceToHost));
        int bad = 0, bad_in_window = 0, field_bad = 0;
        for (int m = 0; m < M; m++)
            for (int n = 0; n < N; n++) {
                const int dA = dot(a, m, b0, n), dB = dot(a, m, b1, n);
                const float z = 12582912.0f + (float)dA + 65536.0f * (float)dB;   // exact in f32 by construction
                uint32_t want;
                memcpy(&want, &z, 4);
                const uint32_t g = got[m * N + n];
                if (g != want) {
                    bad++;
                    if (bad <= 6) printf("   (%d,%d) dA %d dB %d got %08x want %08x\n", m, n, dA, dB, g, want);
                }
                const bool in_window = dB > -64 && dB < 64;
                if (in_window) {
                    if (g != 0x4B400000u + (uint32_t)(dA + 65536 * dB)) bad_in_window++;
                    if ((int16_t)(g & 0xFFFF) != dA) field_bad++;
                }
            }
        printf("A from %s; regions B@%u AL@%u AH@%u AC@%u width %u bias_%s: %d / %d wrong bit patterns; in-window mismatches %d, low-field mismatches %d\n",
               p.a_tmem ? "TMEM" : "smem", p.sf_b, p.sf_al, p.sf_ah, p.sf_ac, p.width, p.bias_last ? "last" : "first", bad, M * N,
               bad_in_window, field_bad);
        printf("   extremes: (5,3) %08x (6,4) %08x (7,9) %08x (8,10) %08x (9,11) %08x (10,12) %08x\n", got[5 * N + 3], got[6 * N + 4],
               got[7 * N + 9], got[8 * N + 10], got[9 * N + 11], got[10 * N + 12]);
        if (bad && p.width == 8 && !p.bias_last) rc = 1;
    }
    return rc;
}
