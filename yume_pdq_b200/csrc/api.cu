// api.cu -- the extern "C" boundary (include/yume_pdq_cuda.h).
//
// Host-side plumbing only: argument checks, host<->device staging (pinned,
// double-buffered for frame batches), per-handle scratch for the matcher, result
// sorting.  All arithmetic on frames / PDQFs / hashes happens in the kernels of
// hasher.cu and matcher.cu; there is no CPU fallback.
#include <cuda_runtime.h>
#include <stdint.h>

#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"

namespace ypdq {

std::atomic<uint64_t> g_launches{0};

int ensure_device(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        (void)cudaGetLastError();
        return YUME_PDQ_CUDA_E_NO_DEVICE;
    }
    if (device < 0 || device >= count || device >= 64)   // per-device tables below and in api.cu hold 64 entries
        return YUME_PDQ_CUDA_E_INVALID;
    static std::atomic<int> checked[64];
    if (checked[device].load() == 0) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
            major != 10) {
            (void)cudaGetLastError();
            return YUME_PDQ_CUDA_E_NO_DEVICE;  // the fatbin holds sm_100a code only
        }
        // keep stream-ordered allocations (the matchers' expanded-needle scratch) cached across
        // synchronisations: with the default release threshold of 0 every cudaMallocAsync after a sync goes
        // back to the driver (measured: +36 ms per 4 MB allocation on a B200 box)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        (void)cudaGetLastError();
        checked[device].store(1);
    }
    YPDQ_CUDA_TRY(cudaSetDevice(device));
    return 0;
}

namespace {

// grow-only device / pinned scratch, one set per (host thread, device, purpose); freed at thread exit
constexpr int kMaxDevices = 64;
struct Scratch {
    void *dev = nullptr;
    size_t dev_bytes = 0;
    void *pin = nullptr;
    size_t pin_bytes = 0;
    int device = -1;
    ~Scratch()
    {
        if (!dev && !pin)
            return;
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) {
            (void)cudaGetLastError();   // runtime already torn down at process exit: nothing left to free
            return;
        }
        if (dev)
            cudaFree(dev);
        if (pin)
            cudaFreeHost(pin);
        if (cur >= 0)
            cudaSetDevice(cur);
        (void)cudaGetLastError();
    }
    // the current device is `device_` (ensure_device ran)
    int reserve(int device_, size_t d, size_t p)
    {
        device = device_;
        if (d > dev_bytes) {
            if (dev)
                cudaFree(dev);
            dev = nullptr;
            dev_bytes = 0;
            cudaError_t e = cudaMalloc(&dev, d);
            if (e != cudaSuccess) {
                (void)cudaGetLastError();
                return e == cudaErrorMemoryAllocation ? YUME_PDQ_CUDA_E_NOMEM : (int)e;
            }
            dev_bytes = d;
        }
        if (p > pin_bytes) {
            if (pin)
                cudaFreeHost(pin);
            pin = nullptr;
            pin_bytes = 0;
            cudaError_t e = cudaMallocHost(&pin, p);
            if (e != cudaSuccess) {
                (void)cudaGetLastError();
                return YUME_PDQ_CUDA_E_NOMEM;
            }
            pin_bytes = p;
        }
        return 0;
    }
};

// A host thread that round-robins over several GPUs keeps one scratch set per device (ensure_device has already
// rejected device >= kMaxDevices... it accepts any index the runtime reports, so clamp here).
Scratch &tls_scratch(int which, int device)
{
    static thread_local Scratch s[3][kMaxDevices];
    return s[which][(unsigned)device % kMaxDevices];
}

// copy stream + events of the chunked host-frame pipeline, per (host thread, device)
struct CopyCtx {
    cudaStream_t stream = nullptr;
    cudaEvent_t copied[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
    int device = -1;
    int init(int device_)
    {
        if (stream)
            return 0;
        device = device_;
        YPDQ_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            YPDQ_CUDA_TRY(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
            YPDQ_CUDA_TRY(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        }
        return 0;
    }
    ~CopyCtx()
    {
        if (!stream)
            return;
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) {
            (void)cudaGetLastError();
            return;
        }
        for (int i = 0; i < 2; i++) {
            if (copied[i]) cudaEventDestroy(copied[i]);
            if (done[i]) cudaEventDestroy(done[i]);
        }
        cudaStreamDestroy(stream);
        if (cur >= 0)
            cudaSetDevice(cur);
        (void)cudaGetLastError();
    }
};
CopyCtx &tls_copy_ctx(int device)
{
    static thread_local CopyCtx c[kMaxDevices];
    return c[(unsigned)device % kMaxDevices];
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Carves device sub-buffers for host-resident arguments out of one Scratch allocation.
struct Stager {
    cudaStream_t s;
    struct Out {
        void *host;
        const void *dev;
        size_t bytes;
    };
    std::vector<Out> outs;
    uint8_t *base = nullptr;
    size_t off = 0;
    bool staged_in = false;   // a host input went through the thread's scratch block
    explicit Stager(cudaStream_t s_) : s(s_) {}
    template <class T>
    int in(const T *p, size_t bytes, const T **dp)
    {
        if (!p || is_device_ptr(p)) {
            *dp = p;
            return 0;
        }
        T *d = reinterpret_cast<T *>(base + off);
        off += align_up(bytes, 256);
        YPDQ_CUDA_TRY(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, s));
        staged_in = true;
        *dp = d;
        return 0;
    }
    template <class T>
    void out(T *p, size_t bytes, T **dp)
    {
        if (!p || is_device_ptr(p)) {
            *dp = p;
            return;
        }
        T *d = reinterpret_cast<T *>(base + off);
        off += align_up(bytes, 256);
        outs.push_back({p, d, bytes});
        *dp = d;
    }
    int finish()
    {
        for (auto &o : outs)
            YPDQ_CUDA_TRY(cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, s));
        // any call with a host pointer returns after completion (yume_pdq_cuda.h): the caller may reuse its
        // buffers, and the thread's scratch block may be overwritten by its next call on another stream
        if (!outs.empty() || staged_in)
            YPDQ_CUDA_TRY(cudaStreamSynchronize(s));
        return 0;
    }
};

template <class T>
inline size_t host_bytes(const T *p, size_t bytes)
{
    return (p && !is_device_ptr(p)) ? align_up(bytes, 256) : 0;
}

// the single-frame latency path's buffers, per (host thread, device): the tent buffer, the finished-CTA counter and the
// frame in device memory, results + staging block in pinned (mapped) host memory
#ifndef YPDQ_ONE_FRAME_PARTS
#define YPDQ_ONE_FRAME_PARTS 4
#endif
// Host copy of a quarter frame into the pinned staging block with NON-TEMPORAL stores: the lines do not stay dirty in the
// core's cache, so the copy engine's reads are not snooped out of it.  Measured on the B200 box (Xeon, one thread,
// profiles/r04/): memcpy 53 us per frame and the call 97 us; non-temporal AVX2 stores 43 us and 76 us; 64-byte AVX-512
// stores the same; memcpy + cldemote 104 / 137 us; write-combined pinned memory changes neither.
__attribute__((target("avx2"))) void stage_copy_nt(float *dst, const float *src, size_t bytes)
{
    const __m256i *s = reinterpret_cast<const __m256i *>(src);
    __m256i *d = reinterpret_cast<__m256i *>(dst);      // 32-byte aligned: page-aligned block, parts of 256 KB
    for (size_t i = 0; i < bytes / 32; i += 4) {
        const __m256i a = _mm256_loadu_si256(s + i), b = _mm256_loadu_si256(s + i + 1), c = _mm256_loadu_si256(s + i + 2),
                      e = _mm256_loadu_si256(s + i + 3);
        _mm256_stream_si256(d + i, a);
        _mm256_stream_si256(d + i + 1, b);
        _mm256_stream_si256(d + i + 2, c);
        _mm256_stream_si256(d + i + 3, e);
    }
    _mm_sfence();
}
void stage_copy(float *dst, const float *src, size_t bytes)
{
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2)
        stage_copy_nt(dst, src, bytes);
    else
        std::memcpy(dst, src, bytes);
}

struct OneFrameCtx {
    int device = -1;
    float *buf1_dev = nullptr;
    unsigned int *done = nullptr;
    OneFrameOut *out = nullptr;
    float *stage = nullptr;
    float *frame_dev = nullptr;
    VisitOut *visit = nullptr;    // yume_pdq_visit_dihedrals_smart_kernel's results (pinned)
    uint32_t seq = 0, visit_seq = 0;
    double stamps_us[8] = {0};    // host-side stamps of the last call (diagnostic)
    int init(int device_)
    {
        if (out)
            return 0;
        device = device_;
        void *d = nullptr, *h = nullptr, *st = nullptr;
        if (cudaMalloc(&d, 128 * 128 * 4 + 256 + 512 * 512 * 4) != cudaSuccess || cudaMemset(d, 0, 128 * 128 * 4 + 256) != cudaSuccess ||
            cudaMallocHost(&h, sizeof(OneFrameOut) + sizeof(VisitOut)) != cudaSuccess ||
            cudaMallocHost(&st, 512 * 512 * 4) != cudaSuccess) {
            (void)cudaGetLastError();
            if (d)
                cudaFree(d);
            if (h)
                cudaFreeHost(h);
            return YUME_PDQ_CUDA_E_NOMEM;
        }
        std::memset(h, 0, sizeof(OneFrameOut) + sizeof(VisitOut));
        buf1_dev = static_cast<float *>(d);
        done = reinterpret_cast<unsigned int *>(static_cast<uint8_t *>(d) + 128 * 128 * 4);
        frame_dev = reinterpret_cast<float *>(static_cast<uint8_t *>(d) + 128 * 128 * 4 + 256);
        out = static_cast<OneFrameOut *>(h);
        visit = reinterpret_cast<VisitOut *>(static_cast<uint8_t *>(h) + sizeof(OneFrameOut));
        stage = static_cast<float *>(st);
        return 0;
    }
    ~OneFrameCtx()
    {
        if (!out)
            return;
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) {
            (void)cudaGetLastError();   // runtime already torn down at process exit
            return;
        }
        cudaFree(buf1_dev);
        cudaFreeHost(out);
        cudaFreeHost(stage);
        if (cur >= 0)
            cudaSetDevice(cur);
        (void)cudaGetLastError();
    }
};
// spin on a sequence number the kernel writes last into mapped host memory (the store is visible ~1 us after it is
// issued; a stream synchronisation wakes up ~8 us later); falls back to the stream when the kernel has ended without it
bool poll_seq(volatile unsigned int *flag, uint32_t seq, cudaStream_t s)
{
    const auto t0 = std::chrono::steady_clock::now();
    uint32_t spins = 0;
    while (*flag != seq) {
        if ((++spins & 0x3FFFu) == 0) {
            if (cudaStreamQuery(s) != cudaErrorNotReady)      // finished (or failed) without the store becoming visible
                break;
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(5))
                break;
        }
    }
    if (*flag != seq) {
        if (cudaStreamSynchronize(s) != cudaSuccess || *flag != seq) {
            (void)cudaGetLastError();
            return false;
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return true;
}

OneFrameCtx &tls_one_frame(int device)
{
    static thread_local OneFrameCtx c[kMaxDevices];
    return c[(unsigned)device % kMaxDevices];
}

int default_device()
{
    const char *e = std::getenv("YUME_PDQ_CUDA_DEVICE");
    return e ? std::atoi(e) : 0;
}

constexpr uint32_t dih(int dx0, int dx1, int dy0, int dy1)
{
    return (uint32_t)(uint8_t)dx0 << 24 | (uint32_t)(uint8_t)dx1 << 16 | (uint32_t)(uint8_t)dy0 << 8 |
           (uint32_t)(uint8_t)dy1;
}

}  // namespace
}  // namespace ypdq

using namespace ypdq;

struct YumePDQ_CudaDb {
    int device = 0;
    int sm_count = 148;
    uint8_t *rows = nullptr;     // ONE allocation holding the two planes: n_rows x 16 B prefixes, then n_rows x 16 B suffixes
    size_t n_rows = 0;
    ypdq::DbPlanes planes() const
    {
        return {reinterpret_cast<const uint4 *>(rows), reinterpret_cast<const uint4 *>(rows) + n_rows};
    }
    uint64_t offset = 0;
    // per-handle scratch for the host-facing calls
    uint8_t *d_needles = nullptr;
    size_t needles_cap = 0;
    YumePDQ_Match *d_out = nullptr;
    size_t out_cap = 0;
    uint32_t *d_count = nullptr;           // [0] count, then 8-byte aligned min key at +8
    unsigned long long *d_min = nullptr;
    void *h_pin = nullptr;                 // pinned: 16 bytes count + min key, then the MatchStats of ALGO_AUTO
    ypdq::MatchStats *stats = nullptr;
    std::mutex mu;
};

extern "C" {

/* ---- (A) reference ABI -------------------------------------------------- */
const uint32_t YUME_PDQ_VERSION_MAJOR = 1;
const uint32_t YUME_PDQ_VERSION_MINOR = 1;
const uint32_t YUME_PDQ_VERSION_PATCH = 0;
const char YUME_PDQ_VERSION_STRING[] = "1.1.0";

/* src/lib.rs:116-133 */
const struct YumePDQ_Dihedrals YUME_PDQ_DIHEDRAL_NORMAL = {dih(1, 0, 0, 1)};
const struct YumePDQ_Dihedrals YUME_PDQ_DIHEDRAL_FLIPPED = {dih(1, 0, 0, -1)};
const struct YumePDQ_Dihedrals YUME_PDQ_DIHEDRAL_FLOPPED = {dih(-1, 0, 0, 1)};
const struct YumePDQ_Dihedrals YUME_PDQ_DIHEDRAL_ROTATED_180 = {dih(-1, 0, 0, -1)};
const struct YumePDQ_Dihedrals YUME_PDQ_DIHEDRAL_ROTATED_90 = {dih(0, 1, -1, 0)};
const struct YumePDQ_Dihedrals YUME_PDQ_DIHEDRAL_ROTATED_270 = {dih(0, 1, 1, 0)};
const struct YumePDQ_Dihedrals YUME_PDQ_DIHEDRAL_ROTATED_90_FLOPPED = {dih(0, -1, -1, 0)};
const struct YumePDQ_Dihedrals YUME_PDQ_DIHEDRAL_FLOPPED_ROTATED_270 = {dih(0, -1, 1, 0)};

const uint32_t YUME_PDQ_CUDA_DIHEDRAL_ORDER[8] = {
    dih(1, 0, 0, 1),  dih(-1, 0, 0, 1), dih(-1, 0, 0, -1), dih(1, 0, 0, -1),
    dih(0, -1, -1, 0), dih(0, 1, 1, 0),  dih(0, -1, 1, 0),  dih(0, 1, -1, 0),
};

float yume_pdq_hash_smart_kernel(const float *input, float *threshold, uint8_t *output, float *buf1,
                                 float *tmp, float *pdqf)
{
    (void)tmp;  // scratch in the reference; not needed here and left untouched
    const float nan = std::numeric_limits<float>::quiet_NaN();
    const int device = default_device();
    if (!input || !output || !buf1 || !pdqf || ensure_device(device) != 0)
        return nan;
    // hash_get_threshold, src/lib.rs:427-455, as a LATENCY path (the reference quotes 81 us per ctypes call,
    // README.md:170-173).  Pageable input: the host moves the frame into pinned memory a quarter at a time (non-temporal
    // stores) and hands every quarter to the copy engine as soon as it is there (the DMA of one overlaps the copy of the
    // next; 2 / 8 / 16 parts measured slower: 81 / 84 / 107 us against 76); pinned
    // input: one DMA; device input: none.  Then ONE launch (hasher.cu hash_one_kernel) that writes {buf1, pdqf,
    // threshold, quality, hash} and finally a sequence number straight into mapped host memory, which this thread
    // polls: no D2H copy, no stream synchronisation.
    OneFrameCtx &cx = tls_one_frame(device);
    if (cx.init(device) != 0)
        return nan;
    cudaStream_t s = nullptr;
    const float *frame = cx.frame_dev;
    const auto tq = std::chrono::steady_clock::now();
    auto stamp = [&](int i) { cx.stamps_us[i] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tq).count(); };
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, input) != cudaSuccess) {
        (void)cudaGetLastError();
        at.type = cudaMemoryTypeUnregistered;
    }
    cudaError_t ce = cudaSuccess;
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) {
        frame = input;
    } else if (at.type == cudaMemoryTypeHost) {
        ce = cudaMemcpyAsync(cx.frame_dev, input, 512 * 512 * sizeof(float), cudaMemcpyHostToDevice, s);
    } else {
        constexpr int parts = YPDQ_ONE_FRAME_PARTS;
        constexpr size_t part = 512 * 512 / parts;      // floats
        for (int c = 0; c < parts && ce == cudaSuccess; c++) {
            stage_copy(cx.stage + c * part, input + c * part, part * sizeof(float));
            ce = cudaMemcpyAsync(cx.frame_dev + c * part, cx.stage + c * part, part * sizeof(float), cudaMemcpyHostToDevice, s);
        }
    }
    if (ce != cudaSuccess) {
        (void)cudaGetLastError();
        return nan;
    }
    stamp(0);
    // README.md:164-168 prints the AVX2 kernel's bisection threshold; opt in with YUME_PDQ_CUDA_THRESHOLD=avx2
    const char *style = std::getenv("YUME_PDQ_CUDA_THRESHOLD");
    const bool avx2_style = style && std::strcmp(style, "avx2") == 0;
    OneFrameOut *h = cx.out;
    const uint32_t seq = ++cx.seq;
    if (launch_hash_one(frame, cx.buf1_dev, cx.done, h, seq, avx2_style, s) != 0)
        return nan;
    stamp(1);
    if (!poll_seq(&h->seq, seq, s))
        return nan;
    stamp(2);
    std::memcpy(buf1, h->buf1, sizeof h->buf1);
    std::memcpy(pdqf, h->pdqf, sizeof h->pdqf);
    std::memcpy(output, h->hash, 32);
    if (threshold)
        *threshold = avx2_style ? h->thr_avx2 : h->thr;
    stamp(3);
    return h->quality;
}

/* diagnostic: host-side microsecond stamps of this thread's last yume_pdq_hash_smart_kernel call -- frame handed to the
 * copy engine, kernel launched, answer seen, outputs copied */
int yume_pdq_cuda_one_frame_last_timing(double *out4, long long *kernel_clocks7)
{
    if (!out4)
        return YUME_PDQ_CUDA_E_INVALID;
    const OneFrameCtx &cx = tls_one_frame(default_device());
    for (int i = 0; i < 4; i++)
        out4[i] = cx.stamps_us[i];
    if (kernel_clocks7)
        for (int i = 0; i < 7; i++)
            kernel_clocks7[i] = cx.out ? cx.out->trace[i + 1] - cx.out->trace[0] : 0;
    return 0;
}

bool yume_pdq_visit_dihedrals_smart_kernel(void *ctx, float threshold, uint8_t *output, float *pdqf,
                                           DihedralCallback callback)
{
    const int device = default_device();
    if (!output || !pdqf || !callback || ensure_device(device) != 0)
        return false;
    // one launch, no copies: the PDQF travels in the launch parameters, the kernel writes the 7 transformed PDQFs, the
    // hashes and the qualities into mapped host memory and then a sequence number this thread polls
    OneFrameCtx &cx = tls_one_frame(device);
    if (cx.init(device) != 0)
        return false;
    cudaStream_t s = nullptr;
    using HostOut = VisitOut;
    HostOut *h = cx.visit;
    VisitIn in;
    std::memcpy(in.p, pdqf, sizeof in.p);
    const uint32_t seq = ++cx.visit_seq;
    if (launch_dihedrals_one(in, threshold, h, seq, s) != 0)
        return false;
    if (!poll_seq(&h->seq, seq, s))
        return false;
    // lib.rs:381-419: 7 callbacks; lib.rs:223-224: the callback sees the ORIGINAL threshold
    HostOut local = *h;  // the callback may re-enter this library on the same thread
    for (int d = 1; d < 8; d++) {
        std::memcpy(pdqf, local.states + (d - 1) * 256, 1024);
        std::memcpy(output, local.hashes + 32 * d, 32);
        if (!callback(ctx, YUME_PDQ_CUDA_DIHEDRAL_ORDER[d], threshold, local.quality[d]))
            return false;
    }
    return true;
}

/* ---- (B) ---------------------------------------------------------------- */
const char *yume_pdq_cuda_error_string(int code)
{
    switch (code) {
    case YUME_PDQ_CUDA_OK: return "ok";
    case YUME_PDQ_CUDA_E_INVALID: return "invalid argument";
    case YUME_PDQ_CUDA_E_NO_DEVICE: return "no usable sm_100 CUDA device (there is no CPU fallback)";
    case YUME_PDQ_CUDA_E_CAPACITY: return "more matches than the output capacity";
    case YUME_PDQ_CUDA_E_TIMEOUT: return "a peer did not answer a mailbox query in time";
    case YUME_PDQ_CUDA_E_NOMEM: return "out of device or pinned memory";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

int yume_pdq_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

const char *yume_pdq_cuda_kernel_ident(void) { return "cuda_sm100a_f32_bitexact_scalar_order"; }
const char *yume_pdq_cuda_matcher_ident(void) { return "CudaMatcher (sm_100a, carry-save prefilter + popc)"; }
uint64_t yume_pdq_cuda_launch_count(void) { return g_launches.load(); }

}  // extern "C"

namespace {

// Hashes n frames into DEVICE output buffers on stream s.  Device-resident frames: one launch.  Host frames:
// double-buffered chunks of `chunk` frames, the H2D copy of chunk i+1 (copy stream) overlaps the kernels of
// chunk i.  after_chunk(first_frame, count) is invoked (host side, work enqueued on s) once the hash
// kernel of each chunk has been enqueued -- yume_pdq_cuda_screen_batch enqueues that chunk's match there.
// d_in: device scratch for 2 * chunk frames (host frames only).
template <class F>
int hash_chunks(int device, const void *frames_v, int fmt, size_t n, uint8_t *d_hashes, float *d_quality,
                float *d_thr, float *d_pdqf, cudaStream_t s, size_t chunk, void *d_in0, F after_chunk)
{
    const uint8_t *frames = static_cast<const uint8_t *>(frames_v);
    const size_t fb = frame_bytes(fmt);
    if (is_device_ptr(frames)) {
        int rc = launch_hash_batch(frames, fmt, n, d_hashes, d_quality, d_thr, d_pdqf, s);
        return rc ? rc : after_chunk((size_t)0, n);
    }
    uint8_t *d_in[2] = {static_cast<uint8_t *>(d_in0), static_cast<uint8_t *>(d_in0) + chunk * fb};
    CopyCtx &cc = tls_copy_ctx(device);
    {
        const int rc0 = cc.init(device);
        if (rc0)
            return rc0;
    }
    cudaStream_t copy_stream = cc.stream;
    cudaEvent_t *ev_copied = cc.copied, *ev_done = cc.done;
    // the copy stream must not start overwriting scratch still in use by earlier work on s
    YPDQ_CUDA_TRY(cudaEventRecord(ev_done[0], s));
    YPDQ_CUDA_TRY(cudaStreamWaitEvent(copy_stream, ev_done[0], 0));
    size_t ci = 0;
    for (size_t f0 = 0; f0 < n; f0 += chunk, ci++) {
        const int b = (int)(ci & 1);
        const size_t cnt = std::min(chunk, n - f0);
        if (ci >= 2)
            YPDQ_CUDA_TRY(cudaStreamWaitEvent(copy_stream, ev_done[b], 0));
        YPDQ_CUDA_TRY(cudaMemcpyAsync(d_in[b], frames + f0 * fb, cnt * fb, cudaMemcpyHostToDevice, copy_stream));
        YPDQ_CUDA_TRY(cudaEventRecord(ev_copied[b], copy_stream));
        YPDQ_CUDA_TRY(cudaStreamWaitEvent(s, ev_copied[b], 0));
        int rc = launch_hash_batch(d_in[b], fmt, cnt, d_hashes + f0 * 256, d_quality ? d_quality + f0 * 8 : nullptr,
                                   d_thr ? d_thr + f0 * 8 : nullptr, d_pdqf ? d_pdqf + f0 * 256 : nullptr, s);
        if (rc)
            return rc;
        YPDQ_CUDA_TRY(cudaEventRecord(ev_done[b], s));   // the frame buffer is free once the hash kernel ran
        rc = after_chunk(f0, cnt);
        if (rc)
            return rc;
    }
    return 0;
}

}  // namespace

extern "C" {

size_t yume_pdq_cuda_frame_bytes(int pixel_format) { return frame_bytes(pixel_format); }

int yume_pdq_cuda_host_alloc(int device, size_t bytes, void **out)
{
    if (!out || bytes == 0)
        return YUME_PDQ_CUDA_E_INVALID;
    *out = nullptr;
    const int rc = ensure_device(device);
    if (rc != 0)
        return rc;
    const cudaError_t e = cudaMallocHost(out, bytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        *out = nullptr;
        return e == cudaErrorMemoryAllocation ? YUME_PDQ_CUDA_E_NOMEM : (int)e;
    }
    return 0;
}

int yume_pdq_cuda_host_free(void *p)
{
    if (!p)
        return 0;
    YPDQ_CUDA_TRY(cudaFreeHost(p));
    return 0;
}

int yume_pdq_cuda_hash_batch(int device, const float *frames, size_t n, uint8_t *hashes,
                             float *quality, float *thresholds, float *pdqf, void *stream)
{
    return yume_pdq_cuda_hash_batch_pixels(device, frames, YUME_PDQ_CUDA_PIXEL_F32, n, hashes, quality,
                                           thresholds, pdqf, stream);
}

int yume_pdq_cuda_hash_batch_pixels(int device, const void *frames, int pixel_format, size_t n,
                                    uint8_t *hashes, float *quality, float *thresholds, float *pdqf,
                                    void *stream)
{
    const size_t fb = frame_bytes(pixel_format);
    if (!fb)
        return YUME_PDQ_CUDA_E_INVALID;
    if (!frames || !hashes)
        return n == 0 ? 0 : YUME_PDQ_CUDA_E_INVALID;
    if (n == 0)
        return 0;
    int rc = ensure_device(device);
    if (rc)
        return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const bool frames_on_host = !is_device_ptr(frames);
    // outputs: staged through one scratch block when they live on the host
    Scratch &so = tls_scratch(0, device);
    size_t out_need = host_bytes(hashes, n * 256) + host_bytes(quality, n * 32) +
                      host_bytes(thresholds, n * 32) + host_bytes(pdqf, n * 1024);
    const size_t chunk = std::min<size_t>(n, 64);  // frames per H2D chunk (64 MiB of f32)
    size_t in_need = frames_on_host ? 2 * chunk * fb : 0;
    rc = so.reserve(device, out_need + in_need + 4096, 0);
    if (rc)
        return rc;
    Stager st(s);
    st.base = reinterpret_cast<uint8_t *>(so.dev);
    uint8_t *d_hashes;
    float *d_quality, *d_thr, *d_pdqf;
    st.out(hashes, n * 256, &d_hashes);
    st.out(quality, n * 32, &d_quality);
    st.out(thresholds, n * 32, &d_thr);
    st.out(pdqf, n * 1024, &d_pdqf);
    void *d_in = st.base + align_up(st.off, 256);
    rc = hash_chunks(device, frames, pixel_format, n, d_hashes, d_quality, d_thr, d_pdqf, s, chunk, d_in,
                     [](size_t, size_t) { return 0; });
    if (rc)
        return rc;
    rc = st.finish();
    if (rc)
        return rc;
    if (frames_on_host)
        YPDQ_CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
}

#define STAGE_BEGIN(device, total_host_bytes)                                    \
    int rc = ensure_device(device);                                              \
    if (rc)                                                                      \
        return rc;                                                               \
    cudaStream_t s = (cudaStream_t)stream;                                       \
    Scratch &so = tls_scratch(0, device);                                        \
    rc = so.reserve(device, (total_host_bytes) + 4096, 0);                       \
    if (rc)                                                                      \
        return rc;                                                               \
    Stager st(s);                                                                \
    st.base = reinterpret_cast<uint8_t *>(so.dev)

int yume_pdq_cuda_jarosz_compress(int device, const float *frames, size_t n, float *buf1, void *stream)
{
    if (n == 0)
        return 0;
    if (!frames || !buf1)
        return YUME_PDQ_CUDA_E_INVALID;
    STAGE_BEGIN(device, host_bytes(frames, n * 1048576) + host_bytes(buf1, n * 65536));
    const float *d_frames;
    float *d_buf1;
    rc = st.in(frames, n * 1048576, &d_frames);
    if (rc)
        return rc;
    st.out(buf1, n * 65536, &d_buf1);
    rc = launch_jarosz(d_frames, n, d_buf1, s);
    return rc ? rc : st.finish();
}

int yume_pdq_cuda_cvt_to_luma8f(int device, const void *pixels, int pixel_format, size_t n_pixels,
                                const float *coeff, float *out, void *stream)
{
    const int fmt = pixel_format & ~YUME_PDQ_CUDA_PIXEL_FMA;
    const size_t bpp = fmt == YUME_PDQ_CUDA_PIXEL_LUMA8 ? 1 : fmt == YUME_PDQ_CUDA_PIXEL_RGB8 ? 3
                       : fmt == YUME_PDQ_CUDA_PIXEL_RGBA8 ? 4 : 0;
    if (!bpp || (fmt == YUME_PDQ_CUDA_PIXEL_LUMA8 && pixel_format != fmt))
        return YUME_PDQ_CUDA_E_INVALID;
    if (n_pixels == 0)
        return 0;
    if (!pixels || !out)
        return YUME_PDQ_CUDA_E_INVALID;
    const uint8_t *px = static_cast<const uint8_t *>(pixels);
    STAGE_BEGIN(device, host_bytes(px, n_pixels * bpp) + host_bytes(out, n_pixels * 4));
    const uint8_t *d_px;
    float *d_out;
    rc = st.in(px, n_pixels * bpp, &d_px);
    if (rc)
        return rc;
    st.out(out, n_pixels * 4, &d_out);
    rc = launch_cvt_luma8f(d_px, pixel_format, n_pixels, coeff, d_out, s);
    return rc ? rc : st.finish();
}

int yume_pdq_cuda_dct2d(int device, const float *buf1, size_t n, float *pdqf, void *stream)
{
    if (n == 0)
        return 0;
    if (!buf1 || !pdqf)
        return YUME_PDQ_CUDA_E_INVALID;
    STAGE_BEGIN(device, host_bytes(buf1, n * 65536) + host_bytes(pdqf, n * 1024));
    const float *d_buf1;
    float *d_pdqf;
    rc = st.in(buf1, n * 65536, &d_buf1);
    if (rc)
        return rc;
    st.out(pdqf, n * 1024, &d_pdqf);
    rc = launch_dct2d(d_buf1, n, d_pdqf, s);
    return rc ? rc : st.finish();
}

int yume_pdq_cuda_sum_of_gradients(int device, const float *pdqf, size_t n, float *gradient,
                                   float *quality, void *stream)
{
    if (n == 0)
        return 0;
    if (!pdqf)
        return YUME_PDQ_CUDA_E_INVALID;
    STAGE_BEGIN(device, host_bytes(pdqf, n * 1024) + host_bytes(gradient, n * 4) + host_bytes(quality, n * 4));
    const float *d_pdqf;
    float *d_g, *d_q;
    rc = st.in(pdqf, n * 1024, &d_pdqf);
    if (rc)
        return rc;
    st.out(gradient, n * 4, &d_g);
    st.out(quality, n * 4, &d_q);
    rc = launch_gradients(d_pdqf, n, d_g, d_q, s);
    return rc ? rc : st.finish();
}

int yume_pdq_cuda_quantize(int device, const float *pdqf, size_t n, float *thresholds, uint8_t *hashes,
                           void *stream)
{
    if (n == 0)
        return 0;
    if (!pdqf || !thresholds || !hashes)
        return YUME_PDQ_CUDA_E_INVALID;
    STAGE_BEGIN(device, host_bytes(pdqf, n * 1024) + host_bytes(thresholds, n * 4) + host_bytes(hashes, n * 32));
    const float *d_pdqf;
    float *d_t;
    uint8_t *d_h;
    rc = st.in(pdqf, n * 1024, &d_pdqf);
    if (rc)
        return rc;
    st.out(thresholds, n * 4, &d_t);
    st.out(hashes, n * 32, &d_h);
    rc = launch_quantize(d_pdqf, n, d_t, d_h, s);
    return rc ? rc : st.finish();
}

int yume_pdq_cuda_threshold_2d_f32(int device, const float *pdqf, size_t n, const float *thresholds,
                                   uint8_t *hashes, void *stream)
{
    if (n == 0)
        return 0;
    if (!pdqf || !thresholds || !hashes)
        return YUME_PDQ_CUDA_E_INVALID;
    STAGE_BEGIN(device, host_bytes(pdqf, n * 1024) + host_bytes(thresholds, n * 4) + host_bytes(hashes, n * 32));
    const float *d_pdqf, *d_t;
    uint8_t *d_h;
    rc = st.in(pdqf, n * 1024, &d_pdqf);
    if (rc)
        return rc;
    rc = st.in(thresholds, n * 4, &d_t);
    if (rc)
        return rc;
    st.out(hashes, n * 32, &d_h);
    rc = launch_threshold(d_pdqf, n, d_t, d_h, s);
    return rc ? rc : st.finish();
}

int yume_pdq_cuda_pdqf_transform(int device, float *pdqf, size_t n, int op, void *stream)
{
    if (n == 0)
        return 0;
    if (!pdqf || op < 0 || op > YUME_PDQ_CUDA_OP_NEGATE_OFF_DIAGONALS)
        return YUME_PDQ_CUDA_E_INVALID;
    STAGE_BEGIN(device, host_bytes(pdqf, n * 1024));
    float *d_p;
    if (is_device_ptr(pdqf)) {
        d_p = pdqf;
    } else {
        const float *tmp;
        rc = st.in((const float *)pdqf, n * 1024, &tmp);
        if (rc)
            return rc;
        d_p = const_cast<float *>(tmp);
        st.outs.push_back({pdqf, d_p, n * 1024});
    }
    rc = launch_transform(d_p, n, op, s);
    return rc ? rc : st.finish();
}

int yume_pdq_cuda_dihedrals_batch(int device, const float *pdqf, const float *thr0, size_t n,
                                  uint8_t *hashes, float *quality, float *thresholds, void *stream)
{
    if (n == 0)
        return 0;
    if (!pdqf || !thr0 || !hashes)
        return YUME_PDQ_CUDA_E_INVALID;
    STAGE_BEGIN(device, host_bytes(pdqf, n * 1024) + host_bytes(thr0, n * 4) + host_bytes(hashes, n * 256) +
                            host_bytes(quality, n * 32) + host_bytes(thresholds, n * 32));
    const float *d_pdqf, *d_t0;
    uint8_t *d_h;
    float *d_q, *d_t;
    rc = st.in(pdqf, n * 1024, &d_pdqf);
    if (rc)
        return rc;
    rc = st.in(thr0, n * 4, &d_t0);
    if (rc)
        return rc;
    st.out(hashes, n * 256, &d_h);
    st.out(quality, n * 32, &d_q);
    st.out(thresholds, n * 32, &d_t);
    rc = launch_dihedrals(d_pdqf, d_t0, n, d_h, d_q, d_t, s);
    return rc ? rc : st.finish();
}

/* ---- matcher ------------------------------------------------------------ */
static int db_alloc(int device, size_t n_rows, uint64_t offset, YumePDQ_CudaDb **out)
{
    if (!out)
        return YUME_PDQ_CUDA_E_INVALID;
    *out = nullptr;
    int rc = ensure_device(device);
    if (rc)
        return rc;
    YumePDQ_CudaDb *db = new (std::nothrow) YumePDQ_CudaDb();
    if (!db)
        return YUME_PDQ_CUDA_E_NOMEM;
    db->device = device;
    db->n_rows = n_rows;
    db->offset = offset;
    cudaDeviceGetAttribute(&db->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaError_t e = cudaMalloc(&db->rows, std::max<size_t>(n_rows * 32, 256));
    if (e == cudaSuccess)
        e = cudaMalloc(&db->d_count, 32);     // count, pad, min key; [4]: match counter of ALGO_AUTO's probes
    if (e == cudaSuccess)
        e = cudaMallocHost(&db->h_pin, 16 + sizeof(ypdq::MatchStats) + 64);
    if (e == cudaSuccess)
        e = cudaMemset(db->d_count, 0, 32);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        yume_pdq_cuda_db_destroy(db);
        return e == cudaErrorMemoryAllocation ? YUME_PDQ_CUDA_E_NOMEM : (int)e;
    }
    db->d_min = reinterpret_cast<unsigned long long *>(db->d_count + 2);
    db->stats = new (static_cast<uint8_t *>(db->h_pin) + 16) ypdq::MatchStats();
    db->stats->probe_count = db->d_count + 4;
    *out = db;
    return 0;
}

// rows (n x 32 B, host or device) -> the two planes of `db`, starting at row `first`
static int split_rows(YumePDQ_CudaDb *db, size_t first, const uint8_t *rows, size_t n)
{
    if (n == 0)
        return 0;
    uint8_t *pre = db->rows + first * 16, *suf = db->rows + (db->n_rows + first) * 16;
    for (size_t r0 = 0; r0 < n; r0 += (size_t)16 << 20) {      // strided copies of at most 16M rows each
        const size_t cnt = std::min<size_t>((size_t)16 << 20, n - r0);
        YPDQ_CUDA_TRY(cudaMemcpy2D(pre + r0 * 16, 16, rows + r0 * 32, 32, 16, cnt, cudaMemcpyDefault));
        YPDQ_CUDA_TRY(cudaMemcpy2D(suf + r0 * 16, 16, rows + r0 * 32 + 16, 32, 16, cnt, cudaMemcpyDefault));
    }
    return 0;
}

int yume_pdq_cuda_db_create(int device, const uint8_t *rows, size_t n_rows, uint64_t global_row_offset,
                            YumePDQ_CudaDb **out)
{
    if (n_rows && !rows)
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = db_alloc(device, n_rows, global_row_offset, out);
    if (rc)
        return rc;
    rc = split_rows(*out, 0, rows, n_rows);
    if (rc) {
        yume_pdq_cuda_db_destroy(*out);
        *out = nullptr;
    }
    return rc;
}

int yume_pdq_cuda_db_create_synthetic(int device, size_t n_rows, uint64_t global_row_offset, uint64_t seed,
                                      YumePDQ_CudaDb **out)
{
    int rc = db_alloc(device, n_rows, global_row_offset, out);
    if (rc)
        return rc;
    // generated in chunks of 32-byte rows (the generator is shared with yume_pdq_cuda_synth_rows), then split
    const size_t chunk = std::min<size_t>(n_rows, (size_t)8 << 20);
    uint8_t *tmp = nullptr;
    if (chunk && cudaMalloc(&tmp, chunk * 32) != cudaSuccess) {
        (void)cudaGetLastError();
        rc = YUME_PDQ_CUDA_E_NOMEM;
    }
    for (size_t r0 = 0; rc == 0 && r0 < n_rows; r0 += chunk) {
        const size_t cnt = std::min(chunk, n_rows - r0);
        rc = launch_synth_rows(tmp, cnt, global_row_offset + r0, seed, nullptr);
        if (rc == 0)
            rc = split_rows(*out, r0, tmp, cnt);
    }
    if (rc == 0) {
        cudaError_t e = cudaDeviceSynchronize();
        rc = e == cudaSuccess ? 0 : (int)e;
    }
    if (tmp)
        cudaFree(tmp);
    if (rc) {
        yume_pdq_cuda_db_destroy(*out);
        *out = nullptr;
    }
    return rc;
}

int yume_pdq_cuda_db_destroy(YumePDQ_CudaDb *db)
{
    if (!db)
        return 0;
    cudaSetDevice(db->device);
    if (db->rows) cudaFree(db->rows);
    if (db->d_needles) cudaFree(db->d_needles);
    if (db->d_out) cudaFree(db->d_out);
    if (db->d_count) cudaFree(db->d_count);
    if (db->h_pin) cudaFreeHost(db->h_pin);
    (void)cudaGetLastError();
    delete db;
    return 0;
}

size_t yume_pdq_cuda_db_len(const YumePDQ_CudaDb *db) { return db ? db->n_rows : 0; }
int yume_pdq_cuda_db_device(const YumePDQ_CudaDb *db) { return db ? db->device : -1; }
int yume_pdq_cuda_db_planes(const YumePDQ_CudaDb *db, const uint8_t **prefix, const uint8_t **suffix)
{
    if (!db || !prefix || !suffix)
        return YUME_PDQ_CUDA_E_INVALID;
    *prefix = db->rows;
    *suffix = db->rows + db->n_rows * 16;
    return 0;
}

int yume_pdq_cuda_db_auto_state(const YumePDQ_CudaDb *db, int *fallback, uint32_t *probe_interval)
{
    if (!db || !db->stats)
        return YUME_PDQ_CUDA_E_INVALID;
    if (fallback)
        *fallback = (int)db->stats->fallback;
    if (probe_interval)
        *probe_interval = db->stats->probe_interval;
    return 0;
}

int yume_pdq_cuda_db_write_rows(YumePDQ_CudaDb *db, size_t first, const uint8_t *rows, size_t n)
{
    if (!db || (n && !rows) || first > db->n_rows || n > db->n_rows - first)
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ensure_device(db->device);
    if (rc)
        return rc;
    return split_rows(db, first, rows, n);
}

// shared host-facing driver: mode LIST (hits copied back + sorted) or MIN
static int match_host(YumePDQ_CudaDb *db, const uint8_t *needles, size_t q, uint32_t threshold, int mode,
                      std::vector<YumePDQ_Match> *hits, size_t want, size_t *total,
                      unsigned long long *min_key, cudaStream_t s)
{
    if (!db || !needles || q == 0 || q > 0xffffffffull)
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ensure_device(db->device);
    if (rc)
        return rc;
    std::lock_guard<std::mutex> lock(db->mu);
    const uint8_t *d_needles = needles;
    if (!is_device_ptr(needles)) {
        if (db->needles_cap < q * 32) {
            if (db->d_needles) cudaFree(db->d_needles);
            db->d_needles = nullptr;
            db->needles_cap = 0;
            YPDQ_CUDA_TRY(cudaMalloc(&db->d_needles, align_up(q * 32, 4096)));
            db->needles_cap = align_up(q * 32, 4096);
        }
        YPDQ_CUDA_TRY(cudaMemcpyAsync(db->d_needles, needles, q * 32, cudaMemcpyHostToDevice, s));
        d_needles = db->d_needles;
    }
    size_t cap = mode == MODE_LIST ? std::max<size_t>(want, 1024) : 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        if (mode == MODE_LIST && db->out_cap < cap) {
            if (db->d_out) cudaFree(db->d_out);
            db->d_out = nullptr;
            db->out_cap = 0;
            cudaError_t e = cudaMalloc(&db->d_out, cap * sizeof(YumePDQ_Match));
            if (e != cudaSuccess) {
                (void)cudaGetLastError();
                return YUME_PDQ_CUDA_E_NOMEM;
            }
            db->out_cap = cap;
        }
        const uint32_t dev_cap = (uint32_t)std::min<size_t>(db->out_cap, 0xffffffffu);
        struct Init {
            uint32_t count, pad;
            unsigned long long min;
        } init = {0, 0, ~0ull};
        std::memcpy(db->h_pin, &init, 16);
        YPDQ_CUDA_TRY(cudaMemcpyAsync(db->d_count, db->h_pin, 16, cudaMemcpyHostToDevice, s));
        rc = launch_match(db->planes(), db->n_rows, db->offset, d_needles, q, threshold, db->d_out,
                          mode == MODE_LIST ? dev_cap : 0, db->d_count, db->d_min, mode,
                          YUME_PDQ_CUDA_ALGO_AUTO, db->sm_count, s, 0, db->stats);
        if (rc)
            return rc;
        YPDQ_CUDA_TRY(cudaMemcpyAsync(db->h_pin, db->d_count, 16, cudaMemcpyDeviceToHost, s));
        YPDQ_CUDA_TRY(cudaStreamSynchronize(s));
        Init res;
        std::memcpy(&res, db->h_pin, 16);
        *total = res.count;
        if (min_key)
            *min_key = res.min;
        if (mode != MODE_LIST)
            return 0;
        if (res.count >= YPDQ_COUNT_SATURATED)   // more than ~4e9 matches: the 32-bit counter stopped counting
            return YUME_PDQ_CUDA_E_CAPACITY;
        if (res.count > dev_cap) {  // rare: grow the device list and rescan so the sorted prefix is exact
            cap = res.count;
            continue;
        }
        try {
            hits->resize(res.count);
        } catch (const std::bad_alloc &) {
            return YUME_PDQ_CUDA_E_NOMEM;
        }
        if (res.count)
            YPDQ_CUDA_TRY(cudaMemcpy(hits->data(), db->d_out, res.count * sizeof(YumePDQ_Match),
                                     cudaMemcpyDeviceToHost));
        std::sort(hits->begin(), hits->end(), [](const YumePDQ_Match &a, const YumePDQ_Match &b) {
            return a.row != b.row ? a.row < b.row : a.needle < b.needle;
        });
        return 0;
    }
    return YUME_PDQ_CUDA_E_CAPACITY;
}

int yume_pdq_cuda_matcher_scan(YumePDQ_CudaDb *db, const uint8_t *needles, size_t q, uint32_t threshold,
                               int *matched, void *stream)
{
    if (!matched)
        return YUME_PDQ_CUDA_E_INVALID;
    size_t total = 0;
    unsigned long long mk;
    int rc = match_host(db, needles, q, threshold, MODE_MIN, nullptr, 0, &total, &mk, (cudaStream_t)stream);
    if (rc)
        return rc;
    *matched = total != 0;
    return 0;
}

int yume_pdq_cuda_matcher_find(YumePDQ_CudaDb *db, const uint8_t *needles, size_t q, uint32_t threshold,
                               YumePDQ_Match *out, size_t cap, size_t *n_found, void *stream)
{
    if (!n_found || (cap && !out))
        return YUME_PDQ_CUDA_E_INVALID;
    std::vector<YumePDQ_Match> hits;
    size_t total = 0;
    int rc = match_host(db, needles, q, threshold, MODE_LIST, &hits, cap, &total, nullptr,
                        (cudaStream_t)stream);
    *n_found = total;     // also on YUME_PDQ_CUDA_E_CAPACITY from a saturated counter (a lower bound then)
    if (rc)
        return rc;
    const size_t k = std::min(cap, hits.size());
    if (k)
        std::memcpy(out, hits.data(), k * sizeof(YumePDQ_Match));
    return total > cap ? YUME_PDQ_CUDA_E_CAPACITY : 0;
}

int yume_pdq_cuda_matcher_find_min(YumePDQ_CudaDb *db, const uint8_t *needles, size_t q, uint32_t threshold,
                                   uint64_t *min_key, void *stream)
{
    if (!min_key)
        return YUME_PDQ_CUDA_E_INVALID;
    size_t total = 0;
    unsigned long long mk = ~0ull;
    int rc = match_host(db, needles, q, threshold, MODE_MIN, nullptr, 0, &total, &mk, (cudaStream_t)stream);
    if (rc)
        return rc;
    *min_key = total ? (uint64_t)mk : UINT64_MAX;
    return 0;
}

int yume_pdq_cuda_matcher_find_device(YumePDQ_CudaDb *db, const uint8_t *needles_dev, size_t q,
                                      uint32_t threshold, YumePDQ_Match *out_dev, uint32_t cap,
                                      uint32_t *count_dev, int algo, void *stream)
{
    if (!db || !needles_dev || !count_dev || q == 0 || (cap && !out_dev))
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ensure_device(db->device);
    if (rc)
        return rc;
    return launch_match(db->planes(), db->n_rows, db->offset, needles_dev, q, threshold, out_dev, cap,
                        count_dev, nullptr, MODE_LIST, algo, db->sm_count, (cudaStream_t)stream, 0, db->stats);
}

int yume_pdq_cuda_matcher_find_device_base(YumePDQ_CudaDb *db, const uint8_t *needles_dev, size_t q,
                                           uint32_t threshold, YumePDQ_Match *out_dev, uint32_t cap,
                                           uint32_t *count_dev, int algo, uint32_t needle_base, void *stream)
{
    if (!db || !needles_dev || !count_dev || q == 0 || (cap && !out_dev))
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ensure_device(db->device);
    if (rc)
        return rc;
    return launch_match(db->planes(), db->n_rows, db->offset, needles_dev, q, threshold, out_dev, cap,
                        count_dev, nullptr, MODE_LIST, algo, db->sm_count, (cudaStream_t)stream, needle_base, db->stats);
}

int yume_pdq_cuda_matcher_find_device_split(YumePDQ_CudaDb *db, const uint8_t *needles_dev, size_t q,
                                            uint32_t threshold, YumePDQ_Match *out_dev, uint32_t cap,
                                            uint32_t *count_dev, int algo, uint32_t needle_base, uint32_t split_at,
                                            uint32_t gap, void *stream)
{
    if (!db || !needles_dev || !count_dev || q == 0 || (cap && !out_dev))
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ensure_device(db->device);
    if (rc)
        return rc;
    return launch_match(db->planes(), db->n_rows, db->offset, needles_dev, q, threshold, out_dev, cap, count_dev, nullptr,
                        MODE_LIST, algo, db->sm_count, (cudaStream_t)stream, NeedleMap(needle_base, split_at, gap), db->stats);
}

int yume_pdq_cuda_screen_batch(YumePDQ_CudaDb *db, const float *frames, size_t n, uint32_t threshold,
                               uint8_t *hashes, float *quality, YumePDQ_Match *out, size_t cap,
                               size_t *n_found, void *stream)
{
    return yume_pdq_cuda_screen_batch_pixels(db, frames, YUME_PDQ_CUDA_PIXEL_F32, n, threshold, hashes, quality,
                                             out, cap, n_found, stream);
}

int yume_pdq_cuda_screen_batch_pixels(YumePDQ_CudaDb *db, const void *frames, int pixel_format, size_t n,
                                      uint32_t threshold, uint8_t *hashes, float *quality,
                                      YumePDQ_Match *out, size_t cap, size_t *n_found, void *stream)
{
    const size_t fb = frame_bytes(pixel_format);
    if (!db || !fb || !n_found || (cap && !out) || (n && !frames) || n > 0x1fffffffull)
        return YUME_PDQ_CUDA_E_INVALID;
    *n_found = 0;
    if (n == 0)
        return 0;
    int rc = ensure_device(db->device);
    if (rc)
        return rc;
    std::lock_guard<std::mutex> lock(db->mu);
    cudaStream_t s = (cudaStream_t)stream;
    const int device = db->device;
    const bool frames_on_host = !is_device_ptr(frames);
    // device staging: hashes always (they are the needles), quality if wanted on the host, frame chunks
    const size_t chunk = std::min<size_t>(n, 128);   // frames per pipeline step: 128 MiB of f32 H2D, 1024 needles
    Scratch &so = tls_scratch(1, device);
    const bool hashes_dev = hashes && is_device_ptr(hashes);
    const size_t need = (hashes_dev ? 0 : align_up(n * 256, 256)) + host_bytes(quality, n * 32) +
                        (frames_on_host ? 2 * chunk * fb : 0) + 4096;
    rc = so.reserve(device, need, 0);
    if (rc)
        return rc;
    uint8_t *base = reinterpret_cast<uint8_t *>(so.dev);
    size_t off = 0;
    uint8_t *d_hashes = hashes_dev ? hashes : base;
    if (!hashes_dev)
        off += align_up(n * 256, 256);
    float *d_quality = quality;
    if (quality && !is_device_ptr(quality)) {
        d_quality = reinterpret_cast<float *>(base + off);
        off += align_up(n * 32, 256);
    }
    void *d_in = base + off;

    size_t list_cap = std::max<size_t>(cap, 1 << 16);
    for (int attempt = 0; attempt < 2; attempt++) {
        if (db->out_cap < list_cap) {
            if (db->d_out) cudaFree(db->d_out);
            db->d_out = nullptr;
            db->out_cap = 0;
            if (cudaMalloc(&db->d_out, list_cap * sizeof(YumePDQ_Match)) != cudaSuccess) {
                (void)cudaGetLastError();
                return YUME_PDQ_CUDA_E_NOMEM;
            }
            db->out_cap = list_cap;
        }
        const uint32_t dev_cap = (uint32_t)std::min<size_t>(db->out_cap, 0xffffffffu);
        struct Init {
            uint32_t count, pad;
            unsigned long long min;
        } init = {0, 0, ~0ull};
        std::memcpy(db->h_pin, &init, 16);
        YPDQ_CUDA_TRY(cudaMemcpyAsync(db->d_count, db->h_pin, 16, cudaMemcpyHostToDevice, s));
        auto match_chunk = [&](size_t f0, size_t cnt) -> int {
            return launch_match(db->planes(), db->n_rows, db->offset, d_hashes + f0 * 256, cnt * 8, threshold, db->d_out,
                                dev_cap, db->d_count, db->d_min, MODE_LIST, YUME_PDQ_CUDA_ALGO_AUTO, db->sm_count, s,
                                (uint32_t)(f0 * 8), db->stats);
        };
        if (attempt == 0) {
            // pipeline: [H2D chunk i+1] || [hash chunk i -> match chunk i's 8 * cnt needles]
            rc = hash_chunks(device, frames, pixel_format, n, d_hashes, d_quality, nullptr, nullptr, s, chunk, d_in,
                             match_chunk);
        } else {
            rc = match_chunk(0, n);   // list overflow: the hashes are resident, rescan with a larger list
        }
        if (rc)
            return rc;
        YPDQ_CUDA_TRY(cudaMemcpyAsync(db->h_pin, db->d_count, 16, cudaMemcpyDeviceToHost, s));
        if (attempt == 0) {
            if (hashes && !hashes_dev)
                YPDQ_CUDA_TRY(cudaMemcpyAsync(hashes, d_hashes, n * 256, cudaMemcpyDeviceToHost, s));
            if (quality && d_quality != quality)
                YPDQ_CUDA_TRY(cudaMemcpyAsync(quality, d_quality, n * 32, cudaMemcpyDeviceToHost, s));
        }
        YPDQ_CUDA_TRY(cudaStreamSynchronize(s));
        Init res;
        std::memcpy(&res, db->h_pin, 16);
        *n_found = res.count;
        if (res.count >= YPDQ_COUNT_SATURATED)   // the 32-bit counter stopped counting: *n_found is a lower bound
            return YUME_PDQ_CUDA_E_CAPACITY;
        if (res.count > dev_cap && attempt == 0) {
            list_cap = res.count;   // the sorted prefix must be exact: grow and rescan
            continue;
        }
        std::vector<YumePDQ_Match> hits;
        try {
            hits.resize(std::min<size_t>(res.count, dev_cap));
        } catch (const std::bad_alloc &) {
            return YUME_PDQ_CUDA_E_NOMEM;
        }
        if (!hits.empty())
            YPDQ_CUDA_TRY(cudaMemcpy(hits.data(), db->d_out, hits.size() * sizeof(YumePDQ_Match),
                                     cudaMemcpyDeviceToHost));
        std::sort(hits.begin(), hits.end(), [](const YumePDQ_Match &a, const YumePDQ_Match &b) {
            return a.row != b.row ? a.row < b.row : a.needle < b.needle;
        });
        const size_t k = std::min(cap, hits.size());
        if (k)
            std::memcpy(out, hits.data(), k * sizeof(YumePDQ_Match));
        return res.count > cap ? YUME_PDQ_CUDA_E_CAPACITY : 0;
    }
    return YUME_PDQ_CUDA_E_CAPACITY;
}

int yume_pdq_cuda_synth_frames(int device, float *frames_dev, size_t n, uint64_t first_frame, uint64_t seed,
                               int kind, void *stream)
{
    if (n && !frames_dev)
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ensure_device(device);
    if (rc)
        return rc;
    return launch_synth_frames(frames_dev, n, first_frame, seed, kind, (cudaStream_t)stream);
}

int yume_pdq_cuda_synth_pixels(int device, void *frames_dev, int pixel_format, size_t n, uint64_t first_frame,
                               uint64_t seed, void *stream)
{
    if (n && !frames_dev)
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ensure_device(device);
    if (rc)
        return rc;
    return launch_synth_pixels(static_cast<uint8_t *>(frames_dev), pixel_format, n, first_frame, seed,
                               (cudaStream_t)stream);
}

int yume_pdq_cuda_synth_rows(int device, uint8_t *rows_dev, size_t n_rows, uint64_t first_row, uint64_t seed,
                             void *stream)
{
    if (n_rows && !rows_dev)
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ensure_device(device);
    if (rc)
        return rc;
    return launch_synth_rows(rows_dev, n_rows, first_row, seed, (cudaStream_t)stream);
}

}  // extern "C"

// ----------------------------------------------------------------------------------------------------------------
// Peer mailboxes: the sharded small-needle query WITHOUT a collective library.  The NCCL formulation of "8 needles on
// one rank -> every rank scans its shard -> did anything match?" costs two collectives and a host synchronisation
// (~95-105 us on 2..8 B200s around a 15-29 us scan).  Here every rank owns a 4 KB mailbox in its HBM that all peers map
// through CUDA IPC, and the exchange is plain stores over NVLink, ordered by flags:
//   root:  post kernel    needles (by value in the launch parameters) -> every rank's mailbox, fence, seq_in = seq
//   all:   wait kernel    spin (bounded) until seq_in == seq, zero the local match counter
//          scan           the usual launch_match on the mailbox's needles, counting only
//          publish kernel count -> root's result[rank], fence, root's seq_out[rank] = seq
//   root:  collect kernel spin until every seq_out == seq, OR the counts into mapped host memory; ONE stream sync.
// The root returns only after it has every rank's answer, so a mailbox is never overwritten while a scan still reads it.
// Where the shard scan is the rows-in-M tensor kernel (matcher_tc1.cu) all five steps ride inside that ONE launch per
// rank (MailboxHook): CTA 0 of the root posts, every CTA waits on its own rank's mailbox, the last CTA to finish
// publishes and, on the root, collects.  Both forms speak the same flag protocol, so ranks may mix them.
namespace ypdq {
namespace {

__global__ void __launch_bounds__(64)
mailbox_post_kernel(MailboxNeedles needles, const uint4 *__restrict__ needles_dev, uint32_t n_vec, MailboxPeers peers, int world,
                    uint32_t seq)
{
    if (threadIdx.x < n_vec) {
        const uint4 v = needles_dev ? needles_dev[threadIdx.x] : needles.v[threadIdx.x];
        for (int r = 0; r < world; r++)
            peers.p[r]->needles[threadIdx.x] = v;
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world)
        st_release_sys(&peers.p[threadIdx.x]->seq_in, seq);
}

__global__ void mailbox_wait_kernel(MailboxMem *mine, uint32_t seq)
{
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(&mine->seq_in) != seq) {
        if (global_ns() - t0 > kMailboxTimeoutNs) {
            mine->error = 1;
            break;
        }
    }
    mine->count[seq & 1u] = 0;
}

__global__ void mailbox_publish_kernel(MailboxMem *mine, MailboxMem *root, int rank, uint32_t seq)
{
    root->result[rank] = mine->count[seq & 1u] | (mine->error ? 0x80000000u : 0u);
    mine->count[(seq + 1u) & 1u] = 0;      // the next query's counter (a fused scan does not run a wait kernel)
    __threadfence_system();
    st_release_sys(&root->seq_out[rank], seq);
}

__global__ void __launch_bounds__(32)
mailbox_collect_kernel(MailboxMem *mine, int world, uint32_t seq, volatile unsigned int *host_out)
{
    mailbox_collect(mine, world, seq, host_out);
}

}  // namespace
}  // namespace ypdq

struct YumePDQ_CudaMailbox {
    int device = 0, rank = 0, world = 1;
    uint32_t seq = 0;
    ypdq::MailboxMem *mine = nullptr;
    ypdq::MailboxPeers peers{};
    unsigned int *h_out = nullptr;     // pinned, mapped: {seq, any mask, error mask}
    bool connected = false;
    int mode = YUME_PDQ_CUDA_MAILBOX_AUTO;
    double timing_us[8] = {0};         // host-side time stamps of the last query on the root (diagnostic)
};

extern "C" {

int yume_pdq_cuda_mailbox_create(int device, int rank, int world, YumePDQ_CudaMailbox **out, uint8_t *handle_out)
{
    using namespace ypdq;
    static_assert(sizeof(cudaIpcMemHandle_t) == YUME_PDQ_CUDA_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    if (!out || !handle_out || world < 1 || world > kMailboxMaxWorld || rank < 0 || rank >= world)
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ensure_device(device);
    if (rc)
        return rc;
    auto *mb = new (std::nothrow) YumePDQ_CudaMailbox();
    if (!mb)
        return YUME_PDQ_CUDA_E_NOMEM;
    mb->device = device;
    mb->rank = rank;
    mb->world = world;
    cudaError_t e = cudaMalloc(&mb->mine, sizeof(MailboxMem));
    if (e == cudaSuccess)
        e = cudaMemset(mb->mine, 0, sizeof(MailboxMem));
    if (e == cudaSuccess)
        e = cudaMallocHost(&mb->h_out, 64);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess)
        e = cudaIpcGetMemHandle(&h, mb->mine);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        yume_pdq_cuda_mailbox_destroy(mb);
        return e == cudaErrorMemoryAllocation ? YUME_PDQ_CUDA_E_NOMEM : (int)e;
    }
    std::memset(mb->h_out, 0, 64);
    std::memcpy(handle_out, &h, sizeof(h));
    mb->peers.p[rank] = mb->mine;
    *out = mb;
    return 0;
}

int yume_pdq_cuda_mailbox_connect(YumePDQ_CudaMailbox *mb, const uint8_t *handles)
{
    if (!mb || !handles || mb->connected)
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ypdq::ensure_device(mb->device);
    if (rc)
        return rc;
    for (int r = 0; r < mb->world; r++) {
        if (r == mb->rank)
            continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * sizeof(h), sizeof(h));
        void *p = nullptr;
        YPDQ_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        mb->peers.p[r] = static_cast<ypdq::MailboxMem *>(p);
    }
    mb->connected = true;
    return 0;
}

int yume_pdq_cuda_mailbox_destroy(YumePDQ_CudaMailbox *mb)
{
    if (!mb)
        return 0;
    (void)cudaSetDevice(mb->device);
    for (int r = 0; r < mb->world; r++)
        if (r != mb->rank && mb->peers.p[r])
            (void)cudaIpcCloseMemHandle(mb->peers.p[r]);
    if (mb->mine)
        cudaFree(mb->mine);
    if (mb->h_out)
        cudaFreeHost(mb->h_out);
    (void)cudaGetLastError();
    delete mb;
    return 0;
}

int yume_pdq_cuda_mailbox_scan(YumePDQ_CudaMailbox *mb, YumePDQ_CudaDb *db, const uint8_t *needles, size_t q,
                               uint32_t threshold, int root, uint32_t *matched_ranks, void *stream)
{
    using namespace ypdq;
    if (!mb || !db || !mb->connected || q == 0 || q > (size_t)kMailboxMaxNeedles || root < 0 || root >= mb->world ||
        db->device != mb->device || (mb->rank == root && !needles))
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ensure_device(mb->device);
    if (rc)
        return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t seq = ++mb->seq;
    const auto tq = std::chrono::steady_clock::now();
    auto stamp = [&](int i) { mb->timing_us[i] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tq).count(); };
    const bool fused = mb->mode == YUME_PDQ_CUDA_MAILBOX_FUSED ||
                       (mb->mode == YUME_PDQ_CUDA_MAILBOX_AUTO && auto_picks_tensor_rows(db->n_rows, q, threshold));
    if (fused && db->n_rows > 0 && threshold <= 128) {
        // ONE launch per rank: the rows-in-M scan posts (root), waits, scans, publishes and collects (root) itself
        MailboxHook hook{};
        hook.needles_dev = nullptr;
        if (mb->rank == root) {
            if (is_device_ptr(needles))
                hook.needles_dev = reinterpret_cast<const uint4 *>(needles);
            else
                std::memcpy(hook.needles.v, needles, q * 32);
        }
        hook.peers = mb->peers;
        hook.mine = mb->mine;
        hook.host_out = mb->h_out;
        hook.world = mb->world;
        hook.rank = mb->rank;
        hook.root = root;
        hook.seq = seq;
        hook.n_vec = (uint32_t)(2 * q);
        rc = launch_match_tensor_rows_mailbox(db->planes(), db->n_rows, db->offset, q, threshold, hook, db->sm_count, s);
        if (rc)
            return rc;
        stamp(0);
        stamp(1);
        stamp(2);
        stamp(3);
    } else {
    if (mb->rank == root) {
        MailboxNeedles arg;
        const uint4 *dev_src = nullptr;
        if (is_device_ptr(needles))
            dev_src = reinterpret_cast<const uint4 *>(needles);
        else
            std::memcpy(arg.v, needles, q * 32);
        mailbox_post_kernel<<<1, 64, 0, s>>>(arg, dev_src, (uint32_t)(2 * q), mb->peers, mb->world, seq);
        YPDQ_LAUNCH_CHECK();
    }
    stamp(0);
    mailbox_wait_kernel<<<1, 1, 0, s>>>(mb->mine, seq);
    YPDQ_LAUNCH_CHECK();
    stamp(1);
    rc = launch_match(db->planes(), db->n_rows, db->offset, reinterpret_cast<const uint8_t *>(mb->mine->needles), q, threshold,
                      nullptr, 0, &mb->mine->count[seq & 1u], nullptr, MODE_LIST, YUME_PDQ_CUDA_ALGO_AUTO, db->sm_count, s, 0, nullptr);
    if (rc)
        return rc;
    stamp(2);
    mailbox_publish_kernel<<<1, 1, 0, s>>>(mb->mine, mb->peers.p[root], mb->rank, seq);
    YPDQ_LAUNCH_CHECK();
    if (mb->rank == root) {
        mailbox_collect_kernel<<<1, 32, 0, s>>>(mb->mine, mb->world, seq, mb->h_out);
        YPDQ_LAUNCH_CHECK();
    }
    stamp(3);
    }
    if (mb->rank != root) {
        // nothing to wait for: this rank's answer travels to the root on the stream; the next query's wait kernel is
        // ordered behind it
        if (matched_ranks)
            *matched_ranks = 0;
        return 0;
    }
    // The collect kernel's last store is the sequence number in mapped host memory: polling it sees the answer ~1 us
    // after the store, a stream synchronisation wakes up ~8 us later.
    volatile unsigned int *h = mb->h_out;
    const auto t0 = std::chrono::steady_clock::now();
    uint32_t spins = 0;
    while (h[0] != seq) {
        if ((++spins & 0xFFFu) == 0) {
            if (cudaStreamQuery(s) != cudaErrorNotReady)      // finished (or failed) without the store becoming visible
                break;
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(5))
                break;
        }
    }
    if (h[0] != seq) {
        YPDQ_CUDA_TRY(cudaStreamSynchronize(s));
        if (h[0] != seq)
            return YUME_PDQ_CUDA_E_TIMEOUT;
    }
    stamp(4);
    if (h[2] != 0)
        return YUME_PDQ_CUDA_E_TIMEOUT;
    if (matched_ranks)
        *matched_ranks = h[1];
    return 0;
}

int yume_pdq_cuda_mailbox_set_mode(YumePDQ_CudaMailbox *mb, int mode)
{
    if (!mb || mode < YUME_PDQ_CUDA_MAILBOX_AUTO || mode > YUME_PDQ_CUDA_MAILBOX_FUSED)
        return YUME_PDQ_CUDA_E_INVALID;
    mb->mode = mode;
    return 0;
}

/* diagnostic: host-side micro-second stamps of the last query on the root -- after the post launch, the wait launch, the
 * scan launch, the publish + collect launches, and when the answer arrived */
int yume_pdq_cuda_mailbox_last_timing(const YumePDQ_CudaMailbox *mb, double *out5)
{
    if (!mb || !out5)
        return YUME_PDQ_CUDA_E_INVALID;
    for (int i = 0; i < 5; i++)
        out5[i] = mb->timing_us[i];
    return 0;
}

}  // extern "C"

// ----------------------------------------------------------------------------------------------------------------
// Peer exchange: the all-gather of every rank's needle slice WITHOUT a kernel.  NCCL's all-gather is a kernel: issued
// beside the persistent tensor matcher it takes SMs first, the matcher's CTAs on those SMs start late and the whole
// launch ends late by the all-gather's duration (measured on 4 B200s, bench.py step_timeline: 0.157 ms of all-gather
// -> own-needle match 2.885 ms instead of 2.736).  Here every rank owns a buffer that the others map through CUDA IPC
// and PUSHES its slice into every peer's buffer with copy-engine copies over NVLink, each followed, in stream order, by
// a 4-byte copy of the sequence number into the peer's flag word for this rank; the consumer's stream waits for the
// world - 1 flags with a one-warp kernel right before the launch that needs the data.  A rank's slice lands at its
// position among the OTHER ranks' slices, so the buffer is the packed batch yume_pdq_cuda_matcher_find_device_split
// takes.  Two buffers alternate by sequence number: a peer may be one exchange ahead (the caller's per-step
// synchronisation -- the match-list gather in ShardedMatcher / bench.py -- keeps it from being two ahead).
namespace ypdq {
namespace {

constexpr int kExchangeRing = 65536;      // ring[i] = i: the source words of the flag copies (16-bit sequence numbers)

struct ExchangeFlags {
    unsigned int from[kMailboxMaxWorld];  // from[r]: sequence number of the last slice rank r has pushed here
    unsigned int error;
    unsigned int pad[15];
};

__global__ void exchange_ring_kernel(unsigned int *ring)
{
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (unsigned)kExchangeRing)
        ring[i] = i;
}

struct ExchangePeers {
    uint4 *dst[kMailboxMaxWorld];            // where this rank's slice goes in every peer's buffer (nullptr: this rank)
    unsigned int *flag[kMailboxMaxWorld];    // the peer's flag word for this rank
};

// The push as ONE kernel: every 16-byte vector of the slice is loaded once and stored to all peers over NVLink; the last
// CTA to finish (device-scope counter, system-scope fences) releases the sequence number into every peer's flag word.
// Issued on the main stream BEFORE the persistent matcher it costs ~15-20 us and then shares the SMs with nobody.
__global__ void __launch_bounds__(256)
exchange_push_kernel(const uint4 *__restrict__ src, size_t n_vec, ExchangePeers peers, int world, unsigned int *done, uint32_t seq)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(src + i);
        for (int p = 0; p < world; p++)
            if (peers.dst[p])
                peers.dst[p][i] = v;
    }
    __threadfence_system();
    __syncthreads();
    __shared__ unsigned int last;
    if (threadIdx.x == 0)
        last = atomicAdd(done, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (!last)
        return;
    if (threadIdx.x == 0)
        *done = 0;
    __threadfence_system();
    if ((int)threadIdx.x < world && peers.flag[threadIdx.x])
        st_release_sys(peers.flag[threadIdx.x], seq);
}

__global__ void __launch_bounds__(32)
exchange_wait_kernel(ExchangeFlags *flags, int world, int rank, uint32_t want)
{
    const int r = threadIdx.x;
    if (r < world && r != rank) {
        const unsigned long long t0 = global_ns();
        // 16-bit sequence numbers: "at or past `want`" is a signed distance
        while ((int16_t)(uint16_t)(ld_acquire_sys(&flags->from[r]) - want) < 0) {
            if (global_ns() - t0 > kMailboxTimeoutNs) {
                flags->error = 1;
                break;
            }
        }
    }
}

}  // namespace
}  // namespace ypdq

struct YumePDQ_CudaExchange {
    int device = 0, rank = 0, world = 1;
    size_t slice_bytes = 0, buf_bytes = 0;      // buf_bytes = (world - 1) * slice_bytes, 256-byte aligned
    uint32_t seq = 0;
    uint8_t *mine = nullptr;                    // [buf 0][buf 1][ExchangeFlags]
    uint8_t *peer[ypdq::kMailboxMaxWorld] = {nullptr};
    unsigned int *ring = nullptr;               // local: [kExchangeRing sequence numbers][finished-CTA counter]
    int sm_count = 148;
    bool connected = false;
    ypdq::ExchangeFlags *flags_of(uint8_t *base) const { return reinterpret_cast<ypdq::ExchangeFlags *>(base + 2 * buf_bytes); }
};

extern "C" {

int yume_pdq_cuda_exchange_create(int device, int rank, int world, size_t slice_bytes, YumePDQ_CudaExchange **out,
                                  uint8_t *handle_out)
{
    using namespace ypdq;
    if (!out || !handle_out || world < 1 || world > kMailboxMaxWorld || rank < 0 || rank >= world || slice_bytes == 0 ||
        slice_bytes % 16 != 0)
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ensure_device(device);
    if (rc)
        return rc;
    auto *ex = new (std::nothrow) YumePDQ_CudaExchange();
    if (!ex)
        return YUME_PDQ_CUDA_E_NOMEM;
    ex->device = device;
    ex->rank = rank;
    ex->world = world;
    ex->slice_bytes = slice_bytes;
    ex->buf_bytes = align_up((size_t)std::max(1, world - 1) * slice_bytes, 256);
    const size_t total = 2 * ex->buf_bytes + sizeof(ExchangeFlags);
    cudaError_t e = cudaMalloc(&ex->mine, total);
    if (e == cudaSuccess)
        e = cudaMemset(ex->mine + 2 * ex->buf_bytes, 0, sizeof(ExchangeFlags));
    if (e == cudaSuccess)
        e = cudaMalloc(&ex->ring, (kExchangeRing + 64) * sizeof(unsigned int));
    if (e == cudaSuccess)
        e = cudaMemset(ex->ring + kExchangeRing, 0, 64 * sizeof(unsigned int));
    if (e == cudaSuccess)
        e = cudaDeviceGetAttribute(&ex->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) {
        exchange_ring_kernel<<<kExchangeRing / 256, 256>>>(ex->ring);
        e = cudaDeviceSynchronize();
    }
    if (e == cudaSuccess)
        e = cudaIpcGetMemHandle(&h, ex->mine);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        yume_pdq_cuda_exchange_destroy(ex);
        return e == cudaErrorMemoryAllocation ? YUME_PDQ_CUDA_E_NOMEM : (int)e;
    }
    std::memcpy(handle_out, &h, sizeof(h));
    ex->peer[rank] = ex->mine;
    *out = ex;
    return 0;
}

int yume_pdq_cuda_exchange_connect(YumePDQ_CudaExchange *ex, const uint8_t *handles)
{
    if (!ex || !handles || ex->connected)
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ypdq::ensure_device(ex->device);
    if (rc)
        return rc;
    for (int r = 0; r < ex->world; r++) {
        if (r == ex->rank)
            continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * sizeof(h), sizeof(h));
        void *p = nullptr;
        YPDQ_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ex->peer[r] = static_cast<uint8_t *>(p);
    }
    ex->connected = true;
    return 0;
}

int yume_pdq_cuda_exchange_post(YumePDQ_CudaExchange *ex, const void *slice_dev, int how, void *stream)
{
    return ex ? yume_pdq_cuda_exchange_post_bytes(ex, slice_dev, ex->slice_bytes, how, stream) : YUME_PDQ_CUDA_E_INVALID;
}

int yume_pdq_cuda_exchange_post_bytes(YumePDQ_CudaExchange *ex, const void *slice_dev, size_t bytes, int how, void *stream)
{
    using namespace ypdq;
    if (!ex || !ex->connected || !slice_dev || (how != YUME_PDQ_CUDA_EXCHANGE_KERNEL && how != YUME_PDQ_CUDA_EXCHANGE_COPY_ENGINE) ||
        bytes == 0 || bytes > ex->slice_bytes || bytes % 16 != 0)
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ensure_device(ex->device);
    if (rc)
        return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t seq = ++ex->seq;
    const size_t buf = (seq & 1u) * ex->buf_bytes;
    if (how == YUME_PDQ_CUDA_EXCHANGE_KERNEL) {
        if (ex->world == 1)
            return 0;
        ExchangePeers peers{};
        for (int p = 0; p < ex->world; p++) {
            if (p == ex->rank)
                continue;
            const size_t slot = (size_t)(ex->rank < p ? ex->rank : ex->rank - 1);
            peers.dst[p] = reinterpret_cast<uint4 *>(ex->peer[p] + buf + slot * bytes);
            peers.flag[p] = &ex->flags_of(ex->peer[p])->from[ex->rank];
        }
        const size_t n_vec = bytes / 16;
        const unsigned grid = (unsigned)std::min<size_t>((n_vec + 255) / 256, (size_t)ex->sm_count * 4);
        exchange_push_kernel<<<grid, 256, 0, s>>>(static_cast<const uint4 *>(slice_dev), n_vec, peers, ex->world,
                                                  ex->ring + kExchangeRing, seq & (kExchangeRing - 1));
        YPDQ_LAUNCH_CHECK();
        return 0;
    }
    // rank r's slice is slot (r < p ? r : r - 1) of peer p's buffer: p's buffer holds every slice but p's own, in rank order
    for (int d = 1; d < ex->world; d++) {
        const int p = (ex->rank + d) % ex->world;      // start with the next rank: the pushes of all ranks fan out evenly
        const size_t slot = (size_t)(ex->rank < p ? ex->rank : ex->rank - 1);
        YPDQ_CUDA_TRY(cudaMemcpyAsync(ex->peer[p] + buf + slot * bytes, slice_dev, bytes, cudaMemcpyDeviceToDevice, s));
    }
    for (int d = 1; d < ex->world; d++) {
        const int p = (ex->rank + d) % ex->world;
        YPDQ_CUDA_TRY(cudaMemcpyAsync(&ex->flags_of(ex->peer[p])->from[ex->rank], ex->ring + (seq & (kExchangeRing - 1)), 4,
                                      cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

int yume_pdq_cuda_exchange_wait(YumePDQ_CudaExchange *ex, void *stream, const uint8_t **others_dev)
{
    using namespace ypdq;
    if (!ex || !ex->connected || !others_dev || ex->seq == 0)
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ensure_device(ex->device);
    if (rc)
        return rc;
    *others_dev = ex->mine + (ex->seq & 1u) * ex->buf_bytes;
    if (ex->world > 1) {
        exchange_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ex->flags_of(ex->mine), ex->world, ex->rank,
                                                               ex->seq & (kExchangeRing - 1));
        YPDQ_LAUNCH_CHECK();
    }
    return 0;
}

int yume_pdq_cuda_exchange_failed(YumePDQ_CudaExchange *ex, int *failed)
{
    if (!ex || !failed)
        return YUME_PDQ_CUDA_E_INVALID;
    int rc = ypdq::ensure_device(ex->device);
    if (rc)
        return rc;
    unsigned int err = 0;
    YPDQ_CUDA_TRY(cudaMemcpy(&err, &ex->flags_of(ex->mine)->error, 4, cudaMemcpyDeviceToHost));
    *failed = err != 0;
    return 0;
}

int yume_pdq_cuda_exchange_destroy(YumePDQ_CudaExchange *ex)
{
    if (!ex)
        return 0;
    (void)cudaSetDevice(ex->device);
    for (int r = 0; r < ex->world; r++)
        if (r != ex->rank && ex->peer[r])
            (void)cudaIpcCloseMemHandle(ex->peer[r]);
    if (ex->mine)
        cudaFree(ex->mine);
    if (ex->ring)
        cudaFree(ex->ring);
    (void)cudaGetLastError();
    delete ex;
    return 0;
}

}  // extern "C"
