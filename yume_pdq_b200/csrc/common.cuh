// common.cuh -- shared helpers for the sm_100a kernels and the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <stdint.h>

#include <atomic>

#include "../../include/yume_pdq_cuda.h"

namespace ypdq {

extern std::atomic<uint64_t> g_launches;  // counted at every kernel launch (api.cu)

#define YPDQ_CUDA_TRY(expr)                                  \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) {                             \
            (void)cudaGetLastError();                        \
            return (int)_e;                                  \
        }                                                    \
    } while (0)

#define YPDQ_LAUNCH_CHECK()                                  \
    do {                                                     \
        ::ypdq::g_launches.fetch_add(1, std::memory_order_relaxed); \
        cudaError_t _e = cudaGetLastError();                 \
        if (_e != cudaSuccess)                               \
            return (int)_e;                                  \
    } while (0)

// true if `p` is memory the device can address directly on `device`
inline bool is_device_ptr(const void *p)
{
    if (!p)
        return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int ensure_device(int device);  // api.cu: validates + cudaSetDevice, returns 0 or error

// ---- kernel launchers (defined in hasher.cu / matcher.cu / synth.cu) ----
int launch_hash_batch(const void *frames, int pixel_format, size_t n, uint8_t *hashes, float *quality,
                      float *thresholds, float *pdqf, cudaStream_t s);
int launch_cvt_luma8f(const void *pixels, int pixel_format, size_t n_pixels, const float *coeff, float *out,
                      cudaStream_t s);
// bytes of one 512 x 512 frame in `pixel_format`, 0 if the format is unknown
inline size_t frame_bytes(int pixel_format)
{
    switch (pixel_format & ~YUME_PDQ_CUDA_PIXEL_FMA) {
    case YUME_PDQ_CUDA_PIXEL_F32: return (pixel_format & YUME_PDQ_CUDA_PIXEL_FMA) ? 0 : (size_t)512 * 512 * 4;
    case YUME_PDQ_CUDA_PIXEL_LUMA8: return (pixel_format & YUME_PDQ_CUDA_PIXEL_FMA) ? 0 : (size_t)512 * 512;
    case YUME_PDQ_CUDA_PIXEL_RGB8: return (size_t)512 * 512 * 3;
    case YUME_PDQ_CUDA_PIXEL_RGBA8: return (size_t)512 * 512 * 4;
    default: return 0;
    }
}
int launch_jarosz(const float *frames, size_t n, float *buf1, cudaStream_t s);
int launch_avx2_style_threshold(const float *pdqf, float *out, cudaStream_t s);
int launch_dct2d(const float *buf1, size_t n, float *pdqf, cudaStream_t s);
int launch_gradients(const float *pdqf, size_t n, float *gradient, float *quality, cudaStream_t s);
int launch_quantize(const float *pdqf, size_t n, float *thresholds, uint8_t *hashes, cudaStream_t s);
int launch_threshold(const float *pdqf, size_t n, const float *thresholds, uint8_t *hashes,
                     cudaStream_t s);
int launch_transform(float *pdqf, size_t n, int op, cudaStream_t s);
int launch_dihedrals(const float *pdqf, const float *thr0, size_t n, uint8_t *hashes, float *quality,
                     float *thresholds, cudaStream_t s);

// The haystack in HBM: two planes of 16 bytes per row -- pre[r] = bytes 0..15 of row r (what every prefix filter
// streams), suf[r] = bytes 16..31 (touched only to verify the rare survivors, or by the unfiltered kernels).  A scan
// that is decided by the prefix moves N x 16 B instead of the N x 32 B an array of 32-byte rows would cost (a 16-byte
// load from a 32-byte row still drags the whole DRAM sector along).
struct DbPlanes {
    const uint4 *pre;
    const uint4 *suf;
};

// What the prefix-filter matcher observed on a haystack: written by the device (pinned, mapped memory) at the end
// of a paired-tensor launch, read by the host when it picks the kernel for the NEXT launch on the same haystack.
// Measured on a B200, 32768 needles x 10M rows (profiles/r02y, r02p):
//  * what the filter pays for is not a candidate but an accumulator BATCH (one epilogue warp's 80 columns x 32 needles
//    = 5120 pairs) that holds at least one and takes the rare path: 5 .. 26 ns of kernel time each (random rows at
//    threshold 36: 2e5 flagged batches, 16.1 ms instead of 10.8; at threshold 40: 5.6e6, 96 ms).  A near-duplicate FAMILY
//    (1024 frames of a video against 10,000 rows near them: 1e7 candidates) flags only 3e5 batches -- all needles of a
//    tile flag the same column at once -- and runs in 11.7 ms;
//  * the unfiltered 256-bit kernel takes twice the time on random rows (14.8e12 pairs/s) and pays ~115 ns per MATCH:
//    780,000 scattered matches take 112 ms where the filter takes 16 ms.
// So the unfiltered kernel is the better choice only when many batches flag AND few of their candidates are matches
// (high thresholds; rows that share a prefix with the needles but not the rest):
//     15 ns x flagged - 115 ns x matches  >  3.3e-5 ns x pairs     and     flagged > 3 % of the batches
// (below ~3 % the pipeline's slack absorbs most of a flagged batch: 1e6 scattered ones cost 5 ms, not 15).
struct MatchStats {
    // device-written (one launch at a time per haystack: the host-facing calls hold the handle's mutex)
    volatile unsigned long long flagged;   // accumulator batches that took the rare path
    volatile unsigned long long matches;   // matches among the listed candidates
    volatile unsigned long long pairs;     // 0: no result since the host last cleared it
    // host-only: adaptive state of ALGO_AUTO
    uint32_t fallback;           // 1: launches run unfiltered; the filter is only probed on a slice of the rows
    uint32_t probe_interval;     // launches between probes while in fallback (8, 16 .. 64)
    uint32_t since_probe;
    uint32_t probe_seq;          // moves the probed slice through the haystack
    uint32_t *probe_count;       // device: match counter of the probes (their matches are discarded)
};
constexpr double kFlaggedBatchNs = 15.0, kUnfilteredMatchNs = 115.0, kUnfilteredExtraNsPerPair = 3.3e-5, kFlaggedBatchFraction = 0.03;

enum MatchMode : int { MODE_LIST = 0, MODE_MIN = 1 };

// The match counter of the C ABI is 32 bits wide.  Reserving slots must never wrap it (32768 needles x 10M rows at a
// threshold near 128 match more than 2^32 pairs): from YPDQ_COUNT_SATURATED on a reservation is taken back, so the
// counter stays put just above that value, the host sees count > capacity and reports YUME_PDQ_CUDA_E_CAPACITY
// (n_found is then a lower bound).  Fewer than 2^28 reservations can be in flight at once, so it cannot wrap.
#define YPDQ_COUNT_SATURATED 0xF0000000u
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t reserve_matches(uint32_t *count, uint32_t n)
{
    const uint32_t pos = atomicAdd(count, n);
    if (pos >= YPDQ_COUNT_SATURATED)
        atomicSub(count, n);
    return pos;
}
#endif
// How a launch's needle index j is reported (LIST mode): base + j, plus `gap` from index `split_at` on -- a batch that is
// a chunk of a longer list (base), possibly the concatenation of two ranges of it with a hole between them (the
// all-gathered needles of every rank but the caller's own)
struct NeedleMap {
    uint32_t base = 0, split_at = 0xFFFFFFFFu, gap = 0;
    NeedleMap() = default;
    NeedleMap(uint32_t base_) : base(base_) {}
    NeedleMap(uint32_t base_, uint32_t split_at_, uint32_t gap_) : base(base_), split_at(split_at_), gap(gap_) {}
#ifdef __CUDACC__
    __host__ __device__
#endif
    uint32_t operator()(uint32_t j) const { return base + j + (j >= split_at ? gap : 0u); }
};

int launch_match(DbPlanes rows, size_t n_rows, uint64_t row_offset, const uint8_t *needles,
                 size_t q, uint32_t threshold, YumePDQ_Match *out, uint32_t cap, uint32_t *count,
                 unsigned long long *min_key, int mode, int algo, int sm_count, cudaStream_t s,
                 NeedleMap needle_base = NeedleMap(), MatchStats *stats = nullptr);

int launch_match_tensor(DbPlanes rows, size_t n_rows, uint64_t row_offset, const uint8_t *needles,
                        size_t q, uint32_t threshold, YumePDQ_Match *out, uint32_t cap, uint32_t *count,
                        unsigned long long *min_key, int mode, int sm_count, cudaStream_t s, NeedleMap needle_base = NeedleMap());

int launch_match_tensor4(DbPlanes rows, size_t n_rows, uint64_t row_offset, const uint8_t *needles,
                         size_t q, uint32_t threshold, YumePDQ_Match *out, uint32_t cap, uint32_t *count,
                         unsigned long long *min_key, int mode, int sm_count, cudaStream_t s, NeedleMap needle_base = NeedleMap(),
                         int prefix = -1);

int launch_match_tensor4x2(DbPlanes rows, size_t n_rows, uint64_t row_offset, const uint8_t *needles,
                           size_t q, uint32_t threshold, YumePDQ_Match *out, uint32_t cap, uint32_t *count,
                           unsigned long long *min_key, int mode, int sm_count, cudaStream_t s,
                           NeedleMap needle_base = NeedleMap(), MatchStats *stats = nullptr);

// 1 .. 32 needles: database rows in the M dimension, operand A written to tensor memory from registers (matcher_tc1.cu)
int launch_match_tensor_rows(DbPlanes rows, size_t n_rows, uint64_t row_offset, const uint8_t *needles, size_t q,
                             uint32_t threshold, YumePDQ_Match *out, uint32_t cap, uint32_t *count,
                             unsigned long long *min_key, int mode, int sm_count, cudaStream_t s, NeedleMap needle_base = NeedleMap());

// ---- peer mailboxes (api.cu; the rows-in-M scan can carry the whole exchange inside its one launch, matcher_tc1.cu) ----
constexpr int kMailboxMaxWorld = 16;
constexpr int kMailboxMaxNeedles = 32;
constexpr unsigned long long kMailboxTimeoutNs = 2000000000ull;    // a peer that never posts must not hang the GPU

struct MailboxMem {
    uint4 needles[2 * kMailboxMaxNeedles];    // q x 32 B
    unsigned int seq_in;                      // written by the root after the needles
    unsigned int count[2];                    // match counter of the local scan of query seq: count[seq & 1]; the other
                                              // one is zeroed for the next query by whoever publishes this one
    unsigned int error;                       // a wait timed out
    unsigned int done;                        // fused scan: CTAs that have finished (reset by the last one)
    unsigned int pad[3];
    unsigned int seq_out[kMailboxMaxWorld];   // on the root: written by rank r after result[r]
    unsigned int result[kMailboxMaxWorld];
};
struct MailboxNeedles {
    uint4 v[2 * kMailboxMaxNeedles];          // 1 KB, by value in the launch parameters: no upload
};
struct MailboxPeers {
    MailboxMem *p[kMailboxMaxWorld];
};
// everything the fused scan needs to post (root), wait, publish and collect (root) inside its own launch
struct MailboxHook {
    MailboxNeedles needles;                   // root, needles in host memory
    const uint4 *needles_dev;                 // root, needles in device memory (else nullptr)
    MailboxPeers peers;
    MailboxMem *mine;
    volatile unsigned int *host_out;          // root: mapped host memory {seq, mask of matching ranks, mask of failed ranks}
    int world, rank, root;
    uint32_t seq, n_vec;
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// root, one warp: wait (bounded) for every rank's answer of query seq, then {any mask, error mask, seq} to mapped host memory
__device__ __forceinline__ void mailbox_collect(MailboxMem *mine, int world, uint32_t seq, volatile unsigned int *host_out)
{
    const unsigned int lane = threadIdx.x & 31u;
    unsigned int r = 0, err = 0;
    if ((int)lane < world) {
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(&mine->seq_out[lane]) != seq) {
            if (global_ns() - t0 > kMailboxTimeoutNs) {
                err = 1;
                break;
            }
        }
        r = *reinterpret_cast<volatile unsigned int *>(&mine->result[lane]);
    }
    const unsigned int any = __ballot_sync(0xFFFFFFFFu, (r & 0x7FFFFFFFu) != 0);
    const unsigned int bad = __ballot_sync(0xFFFFFFFFu, err || (r & 0x80000000u));
    if (lane == 0) {
        host_out[1] = any;       // bit r: rank r's shard holds a match
        host_out[2] = bad;
        __threadfence_system();
        host_out[0] = seq;
    }
}
#endif

// the rows-in-M scan with the mailbox exchange folded into its launch (q <= 32; see MailboxHook)
int launch_match_tensor_rows_mailbox(DbPlanes rows, size_t n_rows, uint64_t row_offset, size_t q, uint32_t threshold,
                                     const MailboxHook &hook, int sm_count, cudaStream_t s);
// would ALGO_AUTO send this query to the rows-in-M kernel?  (the fused mailbox scan exists for that kernel only)
bool auto_picks_tensor_rows(size_t n_rows, size_t q, uint32_t threshold);

// ---- one frame, latency path (hasher.cu hash_one_kernel; api.cu yume_pdq_hash_smart_kernel) ----
struct OneFrameOut {                      // mapped host memory, written by the kernel
    float buf1[128 * 128];
    float pdqf[256];
    float thr, quality, gradient, thr_avx2;
    uint8_t hash[32];
    unsigned int seq;                     // written last
    unsigned int pad[3];
    long long trace[8];                   // diagnostic, SM clocks of the finishing CTA: start, elected, tent buffer staged,
                                          // dct stage 1, stage 2, median + gradients, hash bits
};
struct VisitIn {                          // the PDQF by value in the launch parameters: no upload
    float p[256];
};
struct VisitOut {                         // mapped host memory, written by the kernel
    float states[7 * 256];                // PDQF of dihedrals 1 .. 7
    float quality[8];
    uint8_t hashes[8 * 32];
    unsigned int seq;                     // written last
    unsigned int pad[7];
};
int launch_dihedrals_one(const VisitIn &in, float thr0, VisitOut *out, uint32_t seq, cudaStream_t s);
int launch_hash_one(const float *frame, float *buf1_dev, unsigned int *done, OneFrameOut *out, uint32_t seq, bool avx2_style,
                    cudaStream_t s);

int launch_synth_frames(float *frames, size_t n, uint64_t first, uint64_t seed, int kind,
                        cudaStream_t s);
int launch_synth_pixels(uint8_t *frames, int pixel_format, size_t n, uint64_t first, uint64_t seed,
                        cudaStream_t s);
int launch_synth_rows(uint8_t *rows, size_t n, uint64_t first, uint64_t seed, cudaStream_t s);

}  // namespace ypdq
