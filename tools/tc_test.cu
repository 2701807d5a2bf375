// tc_test.cu -- bring-up harness for the tcgen05 (+-1 int8) Hamming tile: one CTA computes a
// 128 x 256 x K=256 tile D = A . B^T from host-packed canonical K-major SWIZZLE_NONE tiles and the host
// compares with popcounts.  Developer tool, not part of the product.
// build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tc_test tools/tc_test.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}

constexpr int M = 128, N = 256, K = 256;

__global__ void __launch_bounds__(128, 1)
tile_kernel(const uint8_t *__restrict__ a_tile, const uint8_t *__restrict__ b_tile, int32_t *__restrict__ d_out,
            uint32_t lbo, uint32_t sbo, uint32_t kstep_bytes, uint32_t idesc)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sa = smem;                  // 32 KB
    uint8_t *sb = smem + M * K;          // 64 KB
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + M * K + N * K);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < M * K / 16; i += 128)
        reinterpret_cast<uint4 *>(sa)[i] = reinterpret_cast<const uint4 *>(a_tile)[i];
    for (int i = tid; i < N * K / 16; i += 128)
        reinterpret_cast<uint4 *>(sb)[i] = reinterpret_cast<const uint4 *>(b_tile)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (MMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (tid == 0) {
        for (int k = 0; k < K / 32; k++) {
            uint64_t da = make_desc(smem_u32(sa) + k * kstep_bytes, lbo, sbo);
            uint64_t db = make_desc(smem_u32(sb) + k * kstep_bytes, lbo, sbo);
            uint32_t acc = k > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    // wait for the MMAs
    {
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
        } while (!done);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w reads TMEM lanes 32w..32w+31, 256 columns, 32 at a time
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j++)
            d_out[(size_t)tid * N + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// host: pack rows of bits into the canonical K-major no-swizzle tile of +-1 int8:
// offset(r, k) = (r/8)*sbo + (k/16)*lbo + (r%8)*16 + (k%16)
static void pack(const std::vector<uint8_t> &bits, int rows, uint32_t lbo, uint32_t sbo, std::vector<uint8_t> &tile)
{
    tile.assign((size_t)rows * K, 0);
    for (int r = 0; r < rows; r++)
        for (int k = 0; k < K; k++) {
            int bit = (bits[r * 32 + k / 8] >> (k % 8)) & 1;
            size_t off = (size_t)(r / 8) * sbo + (size_t)(k / 16) * lbo + (r % 8) * 16 + (k % 16);
            tile[off] = bit ? 0x01 : 0xFF;
        }
}

int main()
{
    srand(1);
    std::vector<uint8_t> abits(M * 32), bbits(N * 32);
    for (auto &x : abits) x = rand();
    for (auto &x : bbits) x = rand();
    memcpy(&bbits[5 * 32], &abits[7 * 32], 32);  // plant: row 5 == needle 7
    bbits[5 * 32] ^= 3;
    std::vector<int32_t> want(M * N);
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            int d = 0;
            for (int w = 0; w < 32; w++) d += __builtin_popcount(abits[m * 32 + w] ^ bbits[n * 32 + w]);
            want[m * N + n] = 256 - 2 * d;
        }
    uint8_t *da, *db;
    int32_t *dd;
    CK(cudaMalloc(&da, M * K));
    CK(cudaMalloc(&db, N * K));
    CK(cudaMalloc(&dd, M * N * 4));
    const size_t smem = M * K + N * K + 64;
    CK(cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // instruction descriptor: c=S32 (2<<4), a=INT8 (1<<7), b=INT8 (1<<10), K-major both, N>>3 at 17, M>>4 at 24
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    struct Variant { const char *name; uint32_t lbo, sbo_a, sbo_b, kstep; bool swap; } variants[] = {
        // layout 1: 8-row group holds all 16 K-chunks contiguously: LBO=128, SBO=2048, k-step 256
        {"lbo=128 sbo=2048 (K-chunks contiguous)", 128, 2048, 2048, 256, false},
        // same memory layout but LBO/SBO roles swapped in the descriptor
        {"swapped descriptor fields", 128, 2048, 2048, 256, true},
    };
    int rc = 1;
    for (auto &v : variants) {
        std::vector<uint8_t> ta, tb;
        pack(abits, M, v.lbo, v.sbo_a, ta);
        pack(bbits, N, v.lbo, v.sbo_b, tb);
        CK(cudaMemcpy(da, ta.data(), M * K, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(db, tb.data(), N * K, cudaMemcpyHostToDevice));
        CK(cudaMemset(dd, 0x7f, M * N * 4));
        uint32_t lbo = v.swap ? v.sbo_a : v.lbo, sbo = v.swap ? v.lbo : v.sbo_a;
        tile_kernel<<<1, 128, smem>>>(da, db, dd, lbo, sbo, v.kstep, idesc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: kernel error %s\n", v.name, cudaGetErrorString(e)); return 2; }
        std::vector<int32_t> got(M * N);
        CK(cudaMemcpy(got.data(), dd, M * N * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int i = 0; i < M * N; i++) bad += got[i] != want[i];
        printf("%s: mismatches %d / %d; got[0..3] = %d %d %d %d want %d %d %d %d; planted got %d want %d\n", v.name, bad, M * N,
               got[0], got[1], got[2], got[3], want[0], want[1], want[2], want[3], got[7 * N + 5], want[7 * N + 5]);
        if (bad == 0) { rc = 0; printf("OK with variant: %s\n", v.name); break; }
    }
    return rc;
}
