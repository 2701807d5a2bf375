// x1_test.cu -- bring-up harness (developer tool, not product) for the "rows in the M dimension" tensor scan:
// 128 database rows are the A operand of tcgen05.mma kind::mxf4, written to TENSOR MEMORY straight from registers
// (tcgen05.st, one row per thread = per TMEM lane, masked {0, s} encoding of tc4_common.cuh), the (few) needles are the
// B operand in shared memory (N = 16 / 32), and D[row][needle] = popc(needle) - hamming128(needle, row).
// Checks (1) the TMEM layout of an A operand written with tcgen05.st.32x32b, (2) N = 16 / 32 instruction descriptors,
// (3) the issue rate of back-to-back small-N MMAs with A in TMEM.
// build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/x1_test tools/x1_test.cu
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
__device__ __forceinline__ uint4 mask32(uint32_t w)
{
    return make_uint4(w & 0x11111111u, w & 0x22222222u, w & 0x44444444u, (w >> 1) & 0x44444444u);
}
__device__ __forceinline__ uint4 needle32(uint32_t v)
{
    const uint32_t n = ~v;
    return make_uint4(0x44444444u | ((n & 0x11111111u) << 3), 0x22222222u | ((n & 0x22222222u) << 2),
                      0x11111111u | ((n & 0x44444444u) << 1), 0x11111111u | (n & 0x88888888u));
}

constexpr int M = 128;

template <int N>
__global__ void __launch_bounds__(128, 1)
x1_kernel(const uint4 *__restrict__ rows, const uint4 *__restrict__ needles, float *__restrict__ d_out, long long *clk_out, int reps, int commit_each, int issuers)
{
    __shared__ __align__(1024) uint8_t sB[N * 64];
    __shared__ uint64_t bar;
    __shared__ uint64_t bars8[8];
    __shared__ uint64_t bar4[4];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t IDESC = (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((uint32_t)(M >> 4) << 24);
    if (tid < N) {
        const uint4 v = needles[tid];
        uint8_t *grp = sB + (tid >> 3) * 512 + (tid & 7) * 16;
        *reinterpret_cast<uint4 *>(grp) = needle32(v.x);
        *reinterpret_cast<uint4 *>(grp + 128) = needle32(v.y);
        *reinterpret_cast<uint4 *>(grp + 256) = needle32(v.z);
        *reinterpret_cast<uint4 *>(grp + 384) = needle32(v.w);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        for (int i = 0; i < 8; i++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars8[i])));
        for (int i = 0; i < 4; i++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar4[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t lane_base = tmem + (((uint32_t)warp * 32u) << 16);
    // scale factors: 1.0 (0x7F) everywhere in columns 496..511
    for (uint32_t c = 496; c < 512; c++) {
        uint32_t v = 0x7F7F7F7Fu;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_base + c), "r"(v) : "memory");
    }
    // A: this thread's row -> 16 columns at column 64 (two K = 64 halves of 8 columns)
    {
        const uint4 r = rows[tid];
        const uint4 c0 = mask32(r.x), c1 = mask32(r.y), c2 = mask32(r.z), c3 = mask32(r.w);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
                         lane_base + 64),
                     "r"(c0.x), "r"(c0.y), "r"(c0.z), "r"(c0.w), "r"(c1.x), "r"(c1.y), "r"(c1.z), "r"(c1.w), "r"(c2.x), "r"(c2.y),
                     "r"(c2.z), "r"(c2.w), "r"(c3.x), "r"(c3.y), "r"(c3.z), "r"(c3.w)
                     : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if ((tid & 31) == 0 && warp < issuers) {
        auto mma = [&](uint32_t d, uint32_t a, uint64_t db, uint32_t acc) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%5], [%6], p;\n\t}" ::"r"(d),
                "r"(a), "l"(db), "r"(IDESC), "r"(acc), "r"(tmem + 496), "r"(tmem + 504)
                : "memory");
        };
        const uint64_t b0 = make_desc(s32(sB), 128, 512), b1 = make_desc(s32(sB) + 256, 128, 512);
        const long long t0 = clock64();
        for (int i = 0; i < reps; i++) {
            // a different accumulator every time, as in the scan (8 slots of N columns from column 128)
            const uint32_t d = tmem + 128 + (uint32_t)((i + 2 * warp) % (N <= 32 ? 8 : 256 / N)) * N;
            mma(d, tmem + 64, b0, 0);
            mma(d, tmem + 72, b1, 1);
            if (commit_each)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bars8[i & 7])) : "memory");
            if (commit_each == 2)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar4[warp])) : "memory");
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(s32(&bar4[warp])), "r"(0u) : "memory");
        } while (!done);
        clk_out[warp] = clock64() - t0;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(lane_base + 128 + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; j++)
            d_out[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N>
static int run()
{
    srand(11 + N);
    std::vector<uint32_t> rows(M * 4), needles(N * 4);
    for (auto &x : rows) x = (uint32_t)rand() ^ ((uint32_t)rand() << 16);
    for (auto &x : needles) x = (uint32_t)rand() ^ ((uint32_t)rand() << 16);
    memcpy(&needles[4 * 3], &rows[4 * 5], 16);          // needle 3 == row 5: dot = popc
    for (int i = 0; i < 4; i++) needles[4 * 2 + i] = ~rows[4 * 7 + i];   // needle 2 == ~row 7: dot = popc(n) - 128
    uint4 *dr, *dn;
    float *dd;
    long long *dc;
    CK(cudaMalloc(&dr, M * 16));
    CK(cudaMalloc(&dn, N * 16));
    CK(cudaMalloc(&dd, M * N * 4));
    CK(cudaMalloc(&dc, 32));
    CK(cudaMemcpy(dr, rows.data(), M * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dn, needles.data(), N * 16, cudaMemcpyHostToDevice));
    int rc = 0;
    for (int mode = 0; mode < 6; mode++) {
        const int reps = mode == 0 ? 1 : 2000, commit_each = mode < 2 ? 0 : 1, issuers = mode < 4 ? 1 : (mode == 4 ? 2 : 4);
        x1_kernel<N><<<1, 128>>>(dr, dn, dd, dc, reps, commit_each, issuers);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return 2; }
        std::vector<float> got(M * N);
        long long clk4[4], clk = 0;
        CK(cudaMemcpy(got.data(), dd, M * N * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(clk4, dc, 32, cudaMemcpyDeviceToHost));
        for (int i = 0; i < issuers; i++) clk = clk4[i] > clk ? clk4[i] : clk;
        int bad = 0;
        for (int m = 0; m < M; m++)
            for (int n = 0; n < N; n++) {
                int pn = 0, ham = 0;
                for (int w = 0; w < 4; w++) {
                    pn += __builtin_popcount(needles[4 * n + w]);
                    ham += __builtin_popcount(needles[4 * n + w] ^ rows[4 * m + w]);
                }
                const float want = (float)(pn - ham);
                if (got[m * N + n] != want) {
                    if (++bad <= 6) printf("   (%d,%d) got %g want %g\n", m, n, got[m * N + n], want);
                }
            }
        printf("N = %d, reps %d, commit_each %d, issuers %d: %d / %d wrong; %lld clk for %d x %d MMA pairs = %.1f clk per 128-row tile\n", N, reps, commit_each,
               issuers, bad, M * N, clk, issuers, reps, (double)clk / reps / issuers);
        if (bad) rc = 1;
    }
    return rc;
}

int main()
{
    int rc = run<16>();
    rc |= run<32>();
    rc |= run<64>();
    rc |= run<128>();
    return rc;
}
