// yume_pdq_cuda.hpp -- header-only C++17 mirror of the reference's plugin surface over the C ABI.
//
// The reference's host language is Rust; no Rust toolchain exists in the build image, so this is the
// compiled-language host layer above include/yume_pdq_cuda.h (the Rust glue a maintainer would add is
// in integration/rust/).  Names and argument meaning follow the reference:
//   ypdq::CudaKernel          trait Kernel        src/kernel/mod.rs:268-497
//   ypdq::hash / hash_get_threshold / hash_float / visit_dihedrals   src/lib.rs:322-479
//   ypdq::Dihedrals           src/lib.rs:85-134
//   ypdq::CudaVectorDatabase  VulkanVectorDatabase  src/matching/vulkan.rs:40-65
//   ypdq::CudaMatcher         trait PDQMatcher    src/matching/mod.rs:50-105
// Errors become ypdq::Error (the Rust side would map them to Result); there is no CPU fallback.
#pragma once
#include <array>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "yume_pdq_cuda.h"

namespace ypdq {

struct Error : std::runtime_error {
    int code;
    Error(int c, const char *what) : std::runtime_error(std::string(what) + ": " + yume_pdq_cuda_error_string(c)), code(c) {}
};
inline void check(int rc, const char *what)
{
    if (rc != 0)
        throw Error(rc, what);
}

using PDQHash = std::array<uint8_t, 32>;      // lib.rs:287 PDQHash<U16> flattened
using PDQHashF = std::array<float, 256>;      // lib.rs:290 PDQHashF<f32, U16> flattened

struct Dihedrals {  // lib.rs:85-134
    uint32_t packed;
    static constexpr Dihedrals from_tuples(std::pair<int8_t, int8_t> dx, std::pair<int8_t, int8_t> dy)
    {
        return {uint32_t(uint8_t(dx.first)) << 24 | uint32_t(uint8_t(dx.second)) << 16 |
                uint32_t(uint8_t(dy.first)) << 8 | uint32_t(uint8_t(dy.second))};
    }
    constexpr std::pair<std::pair<int8_t, int8_t>, std::pair<int8_t, int8_t>> into_tuples() const
    {
        return {{int8_t(packed >> 24), int8_t(packed >> 16)}, {int8_t(packed >> 8), int8_t(packed)}};
    }
};

// Batched CUDA implementation of the Kernel trait.  All pointers may be host or device.
class CudaKernel {
public:
    static constexpr int InputDimension = 512, OutputDimension = 16;
    static constexpr int Buffer1WidthX = 128, Buffer1LengthY = 128;
    explicit CudaKernel(int device = 0, void *stream = nullptr) : device_(device), stream_(stream) {}
    const char *ident() const { return yume_pdq_cuda_kernel_ident(); }
    static float adjust_quality(float gradient) { float s = gradient / 180.0f; return s < 1.0f ? s : 1.0f; }

    void jarosz_compress(const float *frames, size_t n, float *buf1) const
    { check(yume_pdq_cuda_jarosz_compress(device_, frames, n, buf1, stream_), "jarosz_compress"); }
    void dct2d(const float *buf1, size_t n, float *pdqf) const
    { check(yume_pdq_cuda_dct2d(device_, buf1, n, pdqf, stream_), "dct2d"); }
    void sum_of_gradients(const float *pdqf, size_t n, float *gradient, float *quality = nullptr) const
    { check(yume_pdq_cuda_sum_of_gradients(device_, pdqf, n, gradient, quality, stream_), "sum_of_gradients"); }
    void quantize(const float *pdqf, size_t n, float *thresholds, uint8_t *hashes) const
    { check(yume_pdq_cuda_quantize(device_, pdqf, n, thresholds, hashes, stream_), "quantize"); }
    void threshold_2d_f32(const float *pdqf, size_t n, const float *thresholds, uint8_t *hashes) const
    { check(yume_pdq_cuda_threshold_2d_f32(device_, pdqf, n, thresholds, hashes, stream_), "threshold_2d_f32"); }
    void pdqf_t(float *pdqf, size_t n) const { transform(pdqf, n, YUME_PDQ_CUDA_OP_T); }
    template <bool NEGATE> void pdqf_negate_alt_cols(float *pdqf, size_t n) const
    { transform(pdqf, n, NEGATE ? YUME_PDQ_CUDA_OP_NEGATE_ALT_COLS_TRUE : YUME_PDQ_CUDA_OP_NEGATE_ALT_COLS_FALSE); }
    template <bool NEGATE> void pdqf_negate_alt_rows(float *pdqf, size_t n) const
    { transform(pdqf, n, NEGATE ? YUME_PDQ_CUDA_OP_NEGATE_ALT_ROWS_TRUE : YUME_PDQ_CUDA_OP_NEGATE_ALT_ROWS_FALSE); }
    void pdqf_negate_off_diagonals(float *pdqf, size_t n) const { transform(pdqf, n, YUME_PDQ_CUDA_OP_NEGATE_OFF_DIAGONALS); }

    // fused hot path: hash_get_threshold + visit_dihedrals for a batch
    void hash_batch(const float *frames, size_t n, uint8_t *hashes, float *quality = nullptr,
                    float *thresholds = nullptr, float *pdqf = nullptr) const
    { check(yume_pdq_cuda_hash_batch(device_, frames, n, hashes, quality, thresholds, pdqf, stream_), "hash_batch"); }
    // the same for raw 8-bit frames (YUME_PDQ_CUDA_PIXEL_LUMA8 / RGB8 / RGBA8): the CLI's input conversion
    // (main.rs:486-608, Kernel::cvt_rgb8_to_luma8f mod.rs:303-340) fused into the hash kernel
    void hash_batch_pixels(const void *frames, int pixel_format, size_t n, uint8_t *hashes, float *quality = nullptr,
                           float *thresholds = nullptr, float *pdqf = nullptr) const
    { check(yume_pdq_cuda_hash_batch_pixels(device_, frames, pixel_format, n, hashes, quality, thresholds, pdqf, stream_),
            "hash_batch_pixels"); }
    // Kernel::cvt_rgb8_to_luma8f / cvt_rgba8_to_luma8f on their own (coeff = {R, G, B}, nullptr = ITU table)
    void cvt_to_luma8f(const void *pixels, int pixel_format, size_t n_pixels, float *out, const float *coeff = nullptr) const
    { check(yume_pdq_cuda_cvt_to_luma8f(device_, pixels, pixel_format, n_pixels, coeff, out, stream_), "cvt_to_luma8f"); }
    void dihedrals_batch(const float *pdqf, const float *thr0, size_t n, uint8_t *hashes, float *quality,
                         float *thresholds) const
    { check(yume_pdq_cuda_dihedrals_batch(device_, pdqf, thr0, n, hashes, quality, thresholds, stream_), "dihedrals_batch"); }

private:
    void transform(float *pdqf, size_t n, int op) const
    { check(yume_pdq_cuda_pdqf_transform(device_, pdqf, n, op, stream_), "pdqf_transform"); }
    int device_;
    void *stream_;
};

inline CudaKernel smart_kernel(int device = 0) { return CudaKernel(device); }  // mod.rs:96-103

// ---- single-frame drivers over the reference's own two C functions (lib.rs:322-479) ----
inline float hash_get_threshold(const float *input, float &threshold, PDQHash &output, float *buf1 /*128*128*/,
                                float *tmp /*128*/, PDQHashF &pdqf)
{ return yume_pdq_hash_smart_kernel(input, &threshold, output.data(), buf1, tmp, pdqf.data()); }
inline float hash(const float *input, PDQHash &output, float *buf1, float *tmp, PDQHashF &pdqf)
{ return yume_pdq_hash_smart_kernel(input, nullptr, output.data(), buf1, tmp, pdqf.data()); }
inline float hash_float(const float *input, PDQHashF &output, float *buf1, float *tmp)
{ PDQHash scratch; return yume_pdq_hash_smart_kernel(input, nullptr, scratch.data(), buf1, tmp, output.data()); }

// F: bool(Dihedrals, float threshold, float quality, PDQHashF&, PDQHash&) -- false stops (Err in Rust)
template <class F>
bool visit_dihedrals(PDQHashF &pdqf, PDQHash &output, float threshold, F &&f)
{
    struct Ctx { F *f; PDQHashF *p; PDQHash *o; } ctx{&f, &pdqf, &output};
    auto tramp = [](void *c, uint32_t d, float thr, float q) -> bool {
        auto *x = static_cast<Ctx *>(c);
        return (*x->f)(Dihedrals{d}, thr, q, *x->p, *x->o);
    };
    return yume_pdq_visit_dihedrals_smart_kernel(&ctx, threshold, output.data(), pdqf.data(), tramp);
}

// ---- matcher ----
class CudaVectorDatabase {
public:
    CudaVectorDatabase(int device, const uint8_t *rows, size_t n_rows, uint64_t global_row_offset = 0)
    { check(yume_pdq_cuda_db_create(device, rows, n_rows, global_row_offset, &db_), "db_create"); }
    static CudaVectorDatabase synthetic(int device, size_t n_rows, uint64_t offset, uint64_t seed)
    { CudaVectorDatabase d; check(yume_pdq_cuda_db_create_synthetic(device, n_rows, offset, seed, &d.db_), "db_create_synthetic"); return d; }
    CudaVectorDatabase(CudaVectorDatabase &&o) noexcept : db_(o.db_) { o.db_ = nullptr; }
    CudaVectorDatabase(const CudaVectorDatabase &) = delete;
    ~CudaVectorDatabase() { yume_pdq_cuda_db_destroy(db_); }
    size_t len() const { return yume_pdq_cuda_db_len(db_); }
    YumePDQ_CudaDb *raw() const { return db_; }
private:
    CudaVectorDatabase() = default;
    YumePDQ_CudaDb *db_ = nullptr;
};

class CudaMatcher {  // PDQMatcher: NoFalseNegative = NoFalsePositive = FindsAllMatches = true
public:
    static constexpr int InputDimension = 32;
    CudaMatcher(const CudaVectorDatabase &haystack, uint32_t threshold, void *stream = nullptr)
        : db_(haystack.raw()), threshold_(threshold), stream_(stream) {}
    static const char *ident() { return yume_pdq_cuda_matcher_ident(); }
    bool scan(const uint8_t *needles, size_t q) const
    { int m = 0; check(yume_pdq_cuda_matcher_scan(db_, needles, q, threshold_, &m, stream_), "scan"); return m != 0; }
    std::vector<YumePDQ_Match> find_all(const uint8_t *needles, size_t q, size_t cap = 1 << 16) const
    {
        for (;;) {
            std::vector<YumePDQ_Match> out(cap);
            size_t n = 0;
            int rc = yume_pdq_cuda_matcher_find(db_, needles, q, threshold_, out.data(), cap, &n, stream_);
            if (rc == YUME_PDQ_CUDA_E_CAPACITY) { cap = n; continue; }
            check(rc, "find");
            out.resize(n);
            return out;
        }
    }
    // hash + 8 dihedrals + match for a batch of frames (examples/end_to_end.rs:224-256 in one pipelined call);
    // needle index of a match = 8 * frame + dihedral
    std::vector<YumePDQ_Match> screen(const void *frames, int pixel_format, size_t n, uint8_t *hashes = nullptr,
                                      float *quality = nullptr, size_t cap = 1 << 16) const
    {
        for (;;) {
            std::vector<YumePDQ_Match> out(cap);
            size_t found = 0;
            int rc = yume_pdq_cuda_screen_batch_pixels(db_, frames, pixel_format, n, threshold_, hashes, quality, out.data(),
                                                       cap, &found, stream_);
            if (rc == YUME_PDQ_CUDA_E_CAPACITY) { cap = found; continue; }
            check(rc, "screen");
            out.resize(found);
            return out;
        }
    }
    // asynchronous device-side scan (device pointers; count_dev is not reset); needle j is reported as needle_base + j,
    // from split_at on as needle_base + gap + j (a batch that is a longer list with a hole cut out)
    void find_device(const uint8_t *needles_dev, size_t q, YumePDQ_Match *out_dev, uint32_t cap, uint32_t *count_dev,
                     uint32_t needle_base = 0, uint32_t split_at = 0xFFFFFFFFu, uint32_t gap = 0,
                     int algo = YUME_PDQ_CUDA_ALGO_AUTO) const
    {
        check(yume_pdq_cuda_matcher_find_device_split(db_, needles_dev, q, threshold_, out_dev, cap, count_dev, algo, needle_base,
                                                      split_at, gap, stream_), "find_device");
    }
    // PDQMatcher::find: f(hay_idx, needle_idx) -> std::optional<R>; first engaged value wins
    template <class R, class F>
    std::optional<R> find(const uint8_t *needles, size_t q, F &&f) const
    {
        for (const auto &m : find_all(needles, q))
            if (auto r = f(size_t(m.row), size_t(m.needle)))
                return r;
        return std::nullopt;
    }
private:
    YumePDQ_CudaDb *db_;
    uint32_t threshold_;
    void *stream_;
};

// ---- one process per GPU: peer-memory mailboxes (the sharded VulkanMatcher::scan) and the needle exchange ----
// `handle()` goes to the peers (any transport), `connect` takes everyone's handles in rank order.
class PeerMailbox {
public:
    PeerMailbox(int device, int rank, int world) { check(yume_pdq_cuda_mailbox_create(device, rank, world, &mb_, handle_), "mailbox_create"); }
    PeerMailbox(const PeerMailbox &) = delete;
    ~PeerMailbox() { yume_pdq_cuda_mailbox_destroy(mb_); }
    const uint8_t *handle() const { return handle_; }
    void connect(const uint8_t *handles) { check(yume_pdq_cuda_mailbox_connect(mb_, handles), "mailbox_connect"); }
    void set_mode(int mode) { check(yume_pdq_cuda_mailbox_set_mode(mb_, mode), "mailbox_set_mode"); }
    // COLLECTIVE; on the root: bit r set = rank r's shard holds a match
    uint32_t scan(const CudaVectorDatabase &shard, const uint8_t *needles, size_t q, uint32_t threshold, int root, void *stream = nullptr)
    {
        uint32_t mask = 0;
        check(yume_pdq_cuda_mailbox_scan(mb_, shard.raw(), needles, q, threshold, root, &mask, stream), "mailbox_scan");
        return mask;
    }
private:
    YumePDQ_CudaMailbox *mb_ = nullptr;
    uint8_t handle_[YUME_PDQ_CUDA_IPC_HANDLE_BYTES];
};

class PeerExchange {
public:
    PeerExchange(int device, int rank, int world, size_t slice_bytes)
    { check(yume_pdq_cuda_exchange_create(device, rank, world, slice_bytes, &ex_, handle_), "exchange_create"); }
    PeerExchange(const PeerExchange &) = delete;
    ~PeerExchange() { yume_pdq_cuda_exchange_destroy(ex_); }
    const uint8_t *handle() const { return handle_; }
    void connect(const uint8_t *handles) { check(yume_pdq_cuda_exchange_connect(ex_, handles), "exchange_connect"); }
    void post(const void *slice_dev, void *stream = nullptr, int how = YUME_PDQ_CUDA_EXCHANGE_KERNEL)
    { check(yume_pdq_cuda_exchange_post(ex_, slice_dev, how, stream), "exchange_post"); }
    // the other ranks' slices of the latest post, packed in rank order (valid until the post after next)
    const uint8_t *wait(void *stream = nullptr)
    {
        const uint8_t *p = nullptr;
        check(yume_pdq_cuda_exchange_wait(ex_, stream, &p), "exchange_wait");
        return p;
    }
    bool failed() { int f = 0; check(yume_pdq_cuda_exchange_failed(ex_, &f), "exchange_failed"); return f != 0; }
private:
    YumePDQ_CudaExchange *ex_ = nullptr;
    uint8_t handle_[YUME_PDQ_CUDA_IPC_HANDLE_BYTES];
};

}  // namespace ypdq
