// mma_peak.cu -- developer microbenchmark (not product): raw tcgen05.mma issue rate on sm_100a for the
// operand kinds a +-1 Hamming GEMM could use: kind::i8 (s8 x s8 -> s32), kind::f8f6f4 (e4m3 -> f32) and
// kind::mxf4.block_scale (packed e2m1, ue8m0 scales = 1.0 -> f32).  One CTA per SM, one thread issues
// ITER x 8 MMAs (M = 128, N = 256, 32 operand bytes of K each) from fixed shared-memory tiles into two
// alternating TMEM accumulators; no loads, no epilogue.  Prints TOPS and SM cycles per MMA.
// build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/mma_peak tools/mma_peak.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (uint64_t)((lbo >> 4) & 0x3FFF) << 16 | (uint64_t)((sbo >> 4) & 0x3FFF) << 32 |
           (uint64_t)1 << 46;
}

constexpr int M = 128, N = 256;
constexpr uint32_t LBO = 128, SBO = 2048;   // 256 bytes of K per row: (r>>3)*2048 + kc*128 + (r&7)*16

template <int KIND>  // 0 i8, 1 f8f6f4 e4m3, 2 mxf4 e2m1 block32
__global__ void __launch_bounds__(128, 1) peak_kernel(int iters, uint32_t idesc, long long *cycles, uint32_t fill)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sa = smem, *sb = smem + M * 256;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + (M + N) * 256);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t x = fill ^ (tid * 2654435761u) ^ blockIdx.x;
    for (int i = tid; i < (M + N) * 256 / 4; i += 128) {
        x = x * 1664525u + 1013904223u;
        uint32_t bits = x >> 8;
        uint32_t w;
        if (KIND == 0) {  // +-1 int8
            w = 0;
            for (int b = 0; b < 4; b++) w |= ((bits >> b) & 1 ? 0x01u : 0xFFu) << (8 * b);
        } else if (KIND == 1) {  // +-1.0 e4m3: 0x38 / 0xB8
            w = 0;
            for (int b = 0; b < 4; b++) w |= ((bits >> b) & 1 ? 0x38u : 0xB8u) << (8 * b);
        } else {  // +-1.0 e2m1 nibbles: 0x2 / 0xA
            w = 0;
            for (int b = 0; b < 8; b++) w |= ((bits >> b) & 1 ? 0x2u : 0xAu) << (4 * b);
        }
        reinterpret_cast<uint32_t *>(smem)[i] = w;
    }
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
    if (KIND == 2) {  // scale factors: ue8m0 127 = 1.0 everywhere in columns 480..511 of every lane
        for (int c = 0; c < 32; c++) {
            uint32_t v = 0x7F7F7F7Fu;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((warp * 32u) << 16) + 480 + c), "r"(v));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t0 = clock64();
    if (tid == 0) {
        for (int it = 0; it < iters; it++) {
            const uint32_t d = tmem + (KIND == 2 ? (it & 1) * 240 : (it & 1) * 256);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint64_t da = make_desc(s32(sa) + k * 2 * LBO, LBO, SBO);
                const uint64_t db = make_desc(s32(sb) + k * 2 * LBO, LBO, SBO);
                const uint32_t acc = k > 0;
                if (KIND == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                else if (KIND == 1)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d),
                                 "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(tmem + 480), "r"(tmem + 496)
                                 : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
    }
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(s32(bar)), "r"(0u) : "memory");
    } while (!done);
    long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int KIND> void run(const char *name, uint32_t idesc, int nmma, int n, int iters)
{
    int sms;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    size_t smem = (M + N) * 256 + 64;
    CK(cudaFuncSetAttribute(peak_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long *cyc;
    CK(cudaMalloc(&cyc, sms * sizeof(long long)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    peak_kernel<KIND><<<sms, 128, smem>>>(iters / 10, idesc, cyc, 1);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    peak_kernel<KIND><<<sms, 128, smem>>>(iters, idesc, cyc, 2);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h[256];
    CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    double total_mma = (double)sms * iters * 8;
    double ops_per_mma = 2.0 * M * n * nmma;
    printf("%-10s %8.3f ms  %8.1f TOPS  %7.1f SM clk per MMA (M128 N%d K%d)  -> %.2f T pairs/s as 256-bit Hamming\n", name, ms,
           total_mma * ops_per_mma / (ms * 1e-3) / 1e12, (double)h[0] / ((double)iters * 8), n, nmma,
           total_mma * (double)M * n * nmma / 256.0 / (ms * 1e-3) / 1e12);
    CK(cudaFree(cyc));
}

int main(int argc, char **argv)
{
    int iters = argc > 1 ? atoi(argv[1]) : 20000;
    const uint32_t n_dim = (uint32_t)(N >> 3) << 17, m_dim = (uint32_t)(M >> 4) << 24;
    run<0>("i8", (2u << 4) | (1u << 7) | (1u << 10) | n_dim | m_dim, 32, N, iters);
    run<1>("f8f6f4", (1u << 4) | n_dim | m_dim, 32, N, iters);
    // mxf4: N = 240 so that two accumulators + scale-factor columns fit in 512 TMEM columns
    const uint32_t n240 = (uint32_t)(240 >> 3) << 17;
    run<2>("mxf4", (1u << 7) | (1u << 10) | n240 | (1u << 23) | m_dim, 64, 240, iters);
    return 0;
}
