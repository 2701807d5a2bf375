// common.cuh -- shared helpers for the sm_100a kernels and the C ABI.
#pragma once
#include <cuda_runtime.h>
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
// Random rows pass the 128/122-bit prefix filter at 2e-8 of the pairs (threshold 31); a database of near-duplicate
// families matched against a video's near-identical frames passes 1e-5 and more, and every candidate costs the
// paired kernel ~13 ns -- past 2.3e-6 the unfiltered 256-bit tensor kernel (14.8e12 pairs/s) is faster.
struct MatchStats {
    // device-written (one launch at a time per haystack: the host-facing calls hold the handle's mutex)
    volatile unsigned long long candidates;
    volatile unsigned long long pairs;
    // host-only: adaptive state of ALGO_AUTO
    uint32_t fallback_left;      // launches to run unfiltered before the filter is probed again
    uint32_t fallback_interval;  // doubles while the probes keep finding a dense haystack (8 .. 256)
};
constexpr double kPrefixFilterBreakEven = 2.3e-6;   // candidates per pair

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
int launch_match(DbPlanes rows, size_t n_rows, uint64_t row_offset, const uint8_t *needles,
                 size_t q, uint32_t threshold, YumePDQ_Match *out, uint32_t cap, uint32_t *count,
                 unsigned long long *min_key, int mode, int algo, int sm_count, cudaStream_t s,
                 uint32_t needle_base = 0, MatchStats *stats = nullptr);

int launch_match_tensor(DbPlanes rows, size_t n_rows, uint64_t row_offset, const uint8_t *needles,
                        size_t q, uint32_t threshold, YumePDQ_Match *out, uint32_t cap, uint32_t *count,
                        unsigned long long *min_key, int mode, int sm_count, cudaStream_t s, uint32_t needle_base = 0);

int launch_match_tensor4(DbPlanes rows, size_t n_rows, uint64_t row_offset, const uint8_t *needles,
                         size_t q, uint32_t threshold, YumePDQ_Match *out, uint32_t cap, uint32_t *count,
                         unsigned long long *min_key, int mode, int sm_count, cudaStream_t s, uint32_t needle_base = 0,
                         int prefix = -1);

int launch_match_tensor4x2(DbPlanes rows, size_t n_rows, uint64_t row_offset, const uint8_t *needles,
                           size_t q, uint32_t threshold, YumePDQ_Match *out, uint32_t cap, uint32_t *count,
                           unsigned long long *min_key, int mode, int sm_count, cudaStream_t s,
                           uint32_t needle_base = 0, MatchStats *stats = nullptr);

int launch_synth_frames(float *frames, size_t n, uint64_t first, uint64_t seed, int kind,
                        cudaStream_t s);
int launch_synth_pixels(uint8_t *frames, int pixel_format, size_t n, uint64_t first, uint64_t seed,
                        cudaStream_t s);
int launch_synth_rows(uint8_t *rows, size_t n, uint64_t first, uint64_t seed, cudaStream_t s);

}  // namespace ypdq
