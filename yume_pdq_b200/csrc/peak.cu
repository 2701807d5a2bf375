// peak.cu -- live roofline denominator for the tensor-core matcher: the raw issue rate of
// tcgen05.mma.kind::mxf4.block_scale (M 128, N 240, K 64, e2m1 +-1.0 operands, unit ue8m0 scales, f32
// accumulate) -- exactly the instruction match_tensor4_kernel issues -- with no loads and no epilogue.
// One CTA per SM; one thread issues `iters` x 8 MMAs from fixed shared-memory tiles into two alternating
// TMEM accumulators, one tcgen05.commit at the end.  bench.py calls this on the box it runs on
// (MEASURED_PEAKS.json has a dense bf16 figure but no FP4 one).
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace ypdq {
namespace {

constexpr int PK_M = 128, PK_N = 240, PK_ROW = 256;       // 256 operand bytes of K per row = 8 MMAs of K = 64 (32 B)
constexpr uint32_t PK_LBO = 128, PK_SBO = 8 * PK_ROW;
constexpr size_t PK_SMEM = (size_t)(PK_M + PK_N) * PK_ROW + 64;
constexpr uint32_t PK_IDESC = (1u << 7) | (1u << 10) | ((uint32_t)(PK_N >> 3) << 17) | (1u << 23) |
                              ((uint32_t)(PK_M >> 4) << 24);

__device__ __forceinline__ uint32_t pk_s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t pk_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (uint64_t)(PK_LBO >> 4) << 16 | (uint64_t)(PK_SBO >> 4) << 32 |
           (uint64_t)1 << 46;
}

__global__ void __launch_bounds__(128, 1) mxf4_peak_kernel(int iters, long long *cycles)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + (PK_M + PK_N) * PK_ROW);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t x = 0x9E3779B9u ^ (tid * 2654435761u) ^ blockIdx.x;
    for (int i = tid; i < (PK_M + PK_N) * PK_ROW / 4; i += 128) {   // random +-1.0 e2m1 nibbles (0x2 / 0xA)
        x = x * 1664525u + 1013904223u;
        uint32_t w = 0;
        for (int b = 0; b < 8; b++)
            w |= (((x >> (8 + b)) & 1u) ? 0x2u : 0xAu) << (4 * b);
        reinterpret_cast<uint32_t *>(smem)[i] = w;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pk_s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pk_s32(slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    for (int c = 0; c < 32; c++)      // ue8m0 127 = 1.0 in every byte of columns 480..511
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((warp * 32u) << 16) + 480 + c),
                     "r"(0x7F7F7F7Fu));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long t0 = clock64();
    if (tid == 0) {
        const uint32_t sa = pk_s32(smem), sb = pk_s32(smem + PK_M * PK_ROW);
        for (int it = 0; it < iters; it++) {
            const uint32_t d = tmem + (it & 1) * PK_N;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint64_t da = pk_desc(sa + k * 2 * PK_LBO), db = pk_desc(sb + k * 2 * PK_LBO);
                const uint32_t acc = k > 0;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d),
                    "l"(da), "l"(db), "r"(PK_IDESC), "r"(acc), "r"(tmem + 480), "r"(tmem + 496)
                    : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(pk_s32(bar))
                     : "memory");
    }
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(pk_s32(bar)), "r"(0u)
            : "memory");
    } while (!done);
    const long long t1 = clock64();
    if (tid == 0)
        cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace
}  // namespace ypdq

using namespace ypdq;

extern "C" int yume_pdq_cuda_measure_mxf4_peak(int device, int iters, double *tera_ops, double *sm_clk_per_mma)
{
    if (iters <= 0 || !tera_ops)
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ensure_device(device);
    if (rc)
        return rc;
    int sms = 0;
    YPDQ_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    YPDQ_CUDA_TRY(cudaFuncSetAttribute(mxf4_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PK_SMEM));
    long long *cyc = nullptr;
    YPDQ_CUDA_TRY(cudaMalloc(&cyc, sms * sizeof(long long)));
    cudaEvent_t e0, e1;
    YPDQ_CUDA_TRY(cudaEventCreate(&e0));
    YPDQ_CUDA_TRY(cudaEventCreate(&e1));
    mxf4_peak_kernel<<<sms, 128, PK_SMEM>>>(iters / 8 + 1, cyc);   // warm-up
    g_launches.fetch_add(1);
    YPDQ_CUDA_TRY(cudaEventRecord(e0));
    mxf4_peak_kernel<<<sms, 128, PK_SMEM>>>(iters, cyc);
    g_launches.fetch_add(1);
    YPDQ_CUDA_TRY(cudaEventRecord(e1));
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0.f;
    long long c0 = 0;
    if (e == cudaSuccess)
        e = cudaEventElapsedTime(&ms, e0, e1);
    if (e == cudaSuccess)
        e = cudaMemcpy(&c0, cyc, sizeof c0, cudaMemcpyDeviceToHost);
    cudaFree(cyc);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return (int)e;
    }
    const double mmas = (double)sms * iters * 8.0;
    *tera_ops = mmas * 2.0 * PK_M * PK_N * 64.0 / (ms * 1e-3) / 1e12;
    if (sm_clk_per_mma)
        *sm_clk_per_mma = (double)c0 / ((double)iters * 8.0);
    return 0;
}
