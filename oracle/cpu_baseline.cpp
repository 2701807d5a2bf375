/*
 * cpu_baseline.cpp -- C++ intrinsics restatement of the reference's FAST CPU
 * path, used only as the timed CPU baseline (bench.py cpu_baseline leg and
 * `bench.py --impl reference`) and cross-checked against pdq_oracle.c in
 * tests/test_cpu_baseline.py.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never linked by the product.
 * The Rust reference cannot be compiled in this image (no rustc), so
 * cpu_baseline.kind is "port".
 *
 * Restated reference code (file:line in /root/reference):
 *   Avx2F32Kernel::jarosz_compress   src/kernel/x86.rs:626-682
 *   Avx2F32Kernel::dct2d_impl        src/kernel/x86.rs:100-281
 *   Avx2F32Kernel::quantize_impl     src/kernel/x86.rs:440-621  (bisection median)
 *   negate rows / cols (AVX2)        src/kernel/x86.rs:285-376
 *   sum_of_gradients, adjust_quality src/kernel/x86.rs:779-785, 841-865
 *   visit_dihedrals schedule         src/lib.rs:358-420
 *   CpuMatcher AVX-512 new/scan/find src/matching/mod.rs:176-203, 260-333, 365-435
 *   CpuMatcher scalar scan/find      src/matching/mod.rs:232-257, 336-362
 * ISA dispatch is at run time (the GPU box's host CPU may differ from the
 * build container's).
 */
#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../include/yume_pdq_luts.h"

#define API extern "C" __attribute__((visibility("default")))
#define AVX2 __attribute__((target("avx2,fma")))
#define AVX512 __attribute__((target("avx512f,avx512vpopcntdq,avx512bw,avx512vl,bmi")))

namespace {

struct Luts {
    alignas(64) float dct[16 * 127 + 16];   // DCT_MATRIX_RMAJOR, 16 zero pads (build.rs:562)
    alignas(32) float tent_x8[7 * 8];       // TENT_FILTER_WEIGHTS_X8 (build.rs:717-748)
    Luts()
    {
        std::memset(dct, 0, sizeof dct);
        std::memset(tent_x8, 0, sizeof tent_x8);
        std::memcpy(dct, YUME_PDQ_DCT_MATRIX_BITS, sizeof(uint32_t) * 16 * 127);
        for (int i = 0; i < 7; i++)
            std::memcpy(tent_x8 + 8 * i, YUME_PDQ_TENT_WEIGHT_BITS + 7 * i, sizeof(uint32_t) * 7);
        // The LUT carries two ~1e-40 (subnormal) residues at [0][4], [0][5] (oracle/lut_gen.py).
        // On Intel hosts every FMA touching them takes a microcode assist and the tent filter
        // runs ~20x slower than the reference publishes (27.6 us, bench.csv:8, measured on
        // Zen 4).  They cannot change any result for non-degenerate images, so the TIMED
        // baseline drops them -- this only makes the CPU baseline faster.
        for (int i = 0; i < 7 * 8; i++)
            if (std::fabs(tent_x8[i]) < 1e-30f)
                tent_x8[i] = 0.0f;
    }
};
const Luts &luts()
{
    static const Luts l;
    return l;
}

// ---- x86.rs:626-682 -------------------------------------------------------
AVX2 void jarosz_compress_avx2(const float *in, float *out /*[128][128]*/)
{
    const Luts &L = luts();
    alignas(32) float ob[8];
    const __m256i shiftl1 = _mm256_set_epi32(7, 7, 6, 5, 4, 3, 2, 1);
    for (int outi = 0; outi < 127; outi++) {
        size_t in_i = (size_t)YUME_PDQ_CONV_OFFSET[outi] - YUME_PDQ_TENT_OFFSET;
        for (int outj = 0; outj < 127; outj++) {
            size_t in_j = (size_t)YUME_PDQ_CONV_OFFSET[outj] - YUME_PDQ_TENT_OFFSET;
            __m256 sum = _mm256_setzero_ps();
            for (int di = 0; di < 7; di++) {
                size_t offset = (in_i + di) * 512 + in_j;
                bool last = (di == 6 && outi == 126 && outj == 126);
                if (last)
                    offset -= 1;
                __m256 b = _mm256_loadu_ps(in + offset);
                if (last)
                    b = _mm256_permutevar8x32_ps(b, shiftl1);
                __m256 w = _mm256_loadu_ps(L.tent_x8 + di * 8);
                sum = _mm256_fmadd_ps(b, w, sum);
            }
            sum = _mm256_hadd_ps(sum, _mm256_setzero_ps());
            sum = _mm256_hadd_ps(sum, _mm256_setzero_ps());
            _mm256_store_ps(ob, sum);
            out[outi * 128 + outj] = ob[0] + ob[4];
        }
    }
}

AVX2 inline float hsum(__m256 a)
{
    const __m256i s1 = _mm256_set_epi32(0, 7, 6, 5, 4, 3, 2, 1);
    const __m256i s2 = _mm256_set_epi32(1, 0, 7, 6, 5, 4, 3, 2);
    const __m256i s4 = _mm256_set_epi32(3, 2, 1, 0, 7, 6, 5, 4);
    __m256 p = _mm256_add_ps(a, _mm256_permutevar8x32_ps(a, s1));
    __m256 q = _mm256_add_ps(p, _mm256_permutevar8x32_ps(p, s2));
    __m256 r = _mm256_add_ps(q, _mm256_permutevar8x32_ps(q, s4));
    return _mm256_cvtss_f32(r);
}
AVX2 inline float hmax(__m256 a)
{
    const __m256i s1 = _mm256_set_epi32(0, 7, 6, 5, 4, 3, 2, 1);
    const __m256i s2 = _mm256_set_epi32(1, 0, 7, 6, 5, 4, 3, 2);
    const __m256i s4 = _mm256_set_epi32(3, 2, 1, 0, 7, 6, 5, 4);
    __m256 p = _mm256_max_ps(a, _mm256_permutevar8x32_ps(a, s1));
    __m256 q = _mm256_max_ps(p, _mm256_permutevar8x32_ps(p, s2));
    __m256 r = _mm256_max_ps(q, _mm256_permutevar8x32_ps(q, s4));
    return _mm256_cvtss_f32(r);
}
AVX2 inline float hmin(__m256 a)
{
    const __m256i s1 = _mm256_set_epi32(0, 7, 6, 5, 4, 3, 2, 1);
    const __m256i s2 = _mm256_set_epi32(1, 0, 7, 6, 5, 4, 3, 2);
    const __m256i s4 = _mm256_set_epi32(3, 2, 1, 0, 7, 6, 5, 4);
    __m256 p = _mm256_min_ps(a, _mm256_permutevar8x32_ps(a, s1));
    __m256 q = _mm256_min_ps(p, _mm256_permutevar8x32_ps(p, s2));
    __m256 r = _mm256_min_ps(q, _mm256_permutevar8x32_ps(q, s4));
    return _mm256_cvtss_f32(r);
}

// ---- x86.rs:100-281 -------------------------------------------------------
// two alternating FMA accumulators in the pattern 1,0,0,1,1,0,0,1 per 8 steps (unrolled so they
// stay in registers, as the reference's macro expansion does)
AVX2 void dct2d_avx2(const float *buf /*[128][128]*/, float *tmp /*[128]*/, float *pdqf)
{
    const float *D = luts().dct;
    for (int k = 0; k < 16; k++) {
        const float *dk = D + k * 127;
        for (int j8 = 0; j8 < 128; j8 += 8) {
            __m256 s0 = _mm256_setzero_ps(), s1 = _mm256_setzero_ps();
            const float *b = buf + j8;
#define STEP(acc, K2) acc = _mm256_fmadd_ps(_mm256_loadu_ps(b + (K2) * 128), _mm256_set1_ps(dk[K2]), acc)
            for (int k2 = 0; k2 < 128; k2 += 8) {  // accumulator pattern 1,0,0,1,1,0,0,1 (x86.rs:170-204)
                STEP(s1, k2 + 0);
                STEP(s0, k2 + 1);
                STEP(s0, k2 + 2);
                STEP(s1, k2 + 3);
                STEP(s1, k2 + 4);
                STEP(s0, k2 + 5);
                STEP(s0, k2 + 6);
                STEP(s1, k2 + 7);
            }
#undef STEP
            __m256 z = _mm256_setzero_ps();
            __m256 a0 = _mm256_blendv_ps(z, s0, _mm256_cmp_ps(s0, z, _CMP_ORD_S));
            __m256 a1 = _mm256_blendv_ps(z, s1, _mm256_cmp_ps(s1, z, _CMP_ORD_S));
            _mm256_storeu_ps(tmp + j8, _mm256_add_ps(a0, a1));
        }
        for (int j = 0; j < 16; j++) {
            __m256 s0 = _mm256_setzero_ps(), s1 = _mm256_setzero_ps();
            const float *dr = D + j * 127;
#define STEP(acc, C) acc = _mm256_fmadd_ps(_mm256_loadu_ps(tmp + 8 * (C)), _mm256_loadu_ps(dr + 8 * (C)), acc)
            for (int c = 0; c < 16; c += 8) {
                STEP(s1, c + 0);
                STEP(s0, c + 1);
                STEP(s0, c + 2);
                STEP(s1, c + 3);
                STEP(s1, c + 4);
                STEP(s0, c + 5);
                STEP(s0, c + 6);
                STEP(s1, c + 7);
            }
#undef STEP
            pdqf[k * 16 + j] = hsum(_mm256_add_ps(s0, s1));
        }
    }
}

// ---- x86.rs:440-621 -------------------------------------------------------
AVX2 void quantize_avx2(const float *in, float *threshold, uint8_t *out)
{
    __m256 max_v = _mm256_set1_ps(-3.40282347e+38f);
    __m256 min_v = _mm256_set1_ps(3.40282347e+38f);
    for (int i = 0; i < 256; i += 8) {
        __m256 row = _mm256_loadu_ps(in + i);
        min_v = _mm256_min_ps(min_v, row);
        max_v = _mm256_max_ps(max_v, row);
    }
    float max = hmax(max_v), min = hmin(min_v);
    const uint32_t half_point = 128;
    float guess = (min + max) * 0.5f;
    uint32_t num_over = 0, num_under = 0;
    for (int iter = 0; iter < 32; iter++) {
        __m256 guess_v = _mm256_set1_ps(guess);
        __m256 rs_lt = _mm256_setzero_ps(), rs_gt = _mm256_setzero_ps();
        __m256 mn = _mm256_set1_ps(min), mx = _mm256_set1_ps(max);
        uint32_t gt_count = 0, lt_count = 0;
        for (int c = 0; c < 32; c++) {
            __m256 row = _mm256_loadu_ps(in + 8 * c);
            __m256 cmp_gt = _mm256_cmp_ps(row, guess_v, _CMP_GT_OQ);
            out[31 - c] = (uint8_t)_mm256_movemask_ps(cmp_gt);
            __m256 cmp_lt = _mm256_cmp_ps(row, guess_v, _CMP_LT_OQ);
            __m256 cmp_max_lt = _mm256_cmp_ps(row, mx, _CMP_LE_OQ);
            __m256 cmp_min_gt = _mm256_cmp_ps(row, mn, _CMP_GE_OQ);
            __m256 mgt = _mm256_and_ps(cmp_gt, cmp_max_lt);
            __m256 mlt = _mm256_and_ps(cmp_lt, cmp_min_gt);
            gt_count += (uint32_t)__builtin_popcount((unsigned)_mm256_movemask_ps(mgt));
            lt_count += (uint32_t)__builtin_popcount((unsigned)_mm256_movemask_ps(mlt));
            rs_gt = _mm256_add_ps(rs_gt, _mm256_and_ps(mgt, row));
            rs_lt = _mm256_add_ps(rs_lt, _mm256_and_ps(mlt, row));
        }
        if (gt_count + num_over > half_point) {
            num_under += lt_count;
            min = guess;
            guess = hsum(rs_gt) / (float)gt_count;
        } else if (lt_count + num_under > half_point) {
            num_over += gt_count;
            max = guess;
            guess = hsum(rs_lt) / (float)lt_count;
        } else {
            break;
        }
    }
    *threshold = guess;
}

// ---- threshold.rs:37-119 --------------------------------------------------
void threshold_2d(const float *p, uint8_t *out, float thr)
{
    for (int byte = 0; byte < 32; byte++) {
        uint8_t w = 0;
        for (int b = 0; b < 8; b++)
            if (p[byte * 8 + b] > thr)
                w |= (uint8_t)(1u << b);
        out[31 - byte] = w;
    }
}

float sum_of_gradients(const float *p)
{
    float g = 0.0f;
    for (int i = 0; i < 15; i++)
        for (int j = 0; j < 16; j++)
            g += std::fabs(p[i * 16 + j] - p[(i + 1) * 16 + j]);
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 15; j++)
            g += std::fabs(p[i * 16 + j] - p[i * 16 + j + 1]);
    return g;
}
inline float adjust_quality(float g)
{
    float s = g / 180.0f;
    return s < 1.0f ? s : 1.0f;
}
// x86.rs:285-326 (NEGATE=true: rows 0,2,..), 330-376 (NEGATE=false: cols 0,2,..)
inline void negate_even_rows(float *p)
{
    for (int i = 0; i < 16; i += 2)
        for (int j = 0; j < 16; j++)
            p[i * 16 + j] *= -1.0f;
}
inline void negate_even_cols(float *p)
{
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 16; j += 2)
            p[i * 16 + j] *= -1.0f;
}
inline void transpose(float *p)
{
    for (int i = 0; i < 16; i++)
        for (int j = i + 1; j < 16; j++)
            std::swap(p[i * 16 + j], p[j * 16 + i]);
}

struct Scratch {
    alignas(64) float buf1[128 * 128];
    alignas(64) float tmp[128];
    Scratch()
    {
        std::memset(buf1, 0, sizeof buf1);
        std::memset(tmp, 0, sizeof tmp);
    }
};

// hash_get_threshold (lib.rs:427-455) + visit_dihedrals (lib.rs:358-420) with the AVX2 kernel
AVX2 void hash_all_avx2(const float *frame, Scratch &s, uint8_t *hashes, float *quality,
                        float *thresholds, float *pdqf_out)
{
    float p[256];
    jarosz_compress_avx2(frame, s.buf1);
    dct2d_avx2(s.buf1, s.tmp, p);
    if (pdqf_out)
        std::memcpy(pdqf_out, p, sizeof p);
    float thr0, thr_c, thr_od, thr_r;
    quality[0] = adjust_quality(sum_of_gradients(p));
    quantize_avx2(p, &thr0, hashes + 0);
    negate_even_cols(p);
    quantize_avx2(p, &thr_c, hashes + 32);
    quality[1] = adjust_quality(sum_of_gradients(p));
    negate_even_rows(p);
    quantize_avx2(p, &thr_od, hashes + 64);
    quality[2] = adjust_quality(sum_of_gradients(p));
    negate_even_cols(p);
    quantize_avx2(p, &thr_r, hashes + 96);
    quality[3] = adjust_quality(sum_of_gradients(p));
    negate_even_rows(p);
    transpose(p);
    threshold_2d(p, hashes + 128, thr0);
    quality[4] = adjust_quality(sum_of_gradients(p));
    negate_even_rows(p);
    threshold_2d(p, hashes + 160, thr_c);
    quality[5] = adjust_quality(sum_of_gradients(p));
    negate_even_cols(p);
    threshold_2d(p, hashes + 192, thr_od);
    quality[6] = adjust_quality(sum_of_gradients(p));
    negate_even_rows(p);
    threshold_2d(p, hashes + 224, thr_r);
    quality[7] = adjust_quality(sum_of_gradients(p));
    const float t[8] = {thr0, thr_c, thr_od, thr_r, thr0, thr_c, thr_od, thr_r};
    std::memcpy(thresholds, t, sizeof t);
}

template <class F>
void parallel_for(size_t n, int threads, F f)
{
    if (threads <= 1 || n <= 1) {
        f(0, n, 0);
        return;
    }
    std::vector<std::thread> pool;
    size_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        size_t lo = std::min(n, per * t), hi = std::min(n, per * (t + 1));
        if (lo >= hi)
            break;
        pool.emplace_back([=] { f(lo, hi, t); });
    }
    for (auto &th : pool)
        th.join();
}

// ---------------------------------------------------------------- matcher --
struct Match {
    uint64_t row;
    uint32_t needle;
    uint32_t dist;
};

// CpuMatcher<U2048, N> AVX-512: N blocks of 8 needles, complemented and pivoted
// (matching/mod.rs:176-203).  find: cmpge(256-dist, 256-threshold) + tzcnt loop
// (365-435); here the callback records every hit (FindsAllMatches = B1).
struct Avx512Matcher {
    int nblocks;
    alignas(64) uint64_t comp[8][4][8];
    uint64_t threshold;
};

AVX512 void avx512_build(Avx512Matcher &m, const uint8_t *needles, int n_needles, uint32_t thr)
{
    m.nblocks = (n_needles + 7) / 8;
    m.threshold = thr;
    for (int b = 0; b < m.nblocks; b++)
        for (int reg = 0; reg < 4; reg++)
            for (int i = 0; i < 8; i++) {
                int idx = b * 8 + i;
                uint64_t w = 0; // padding needles: all-ones complement -> distance 256 - popc(q) ... handled below
                if (idx < n_needles)
                    std::memcpy(&w, needles + 32 * idx + 8 * reg, 8);
                m.comp[b][reg][i] = ~w;
            }
}

template <bool FIND>
AVX512 bool avx512_run(const Avx512Matcher &m, int n_valid, const uint8_t *rows, size_t n_rows,
                       size_t row_base, uint32_t needle_base, std::vector<Match> *hits)
{
    const __m512i thr = _mm512_set1_epi64((long long)(256 - m.threshold));
    const __m512i addend = _mm512_set1_epi64((long long)m.threshold);
    __m512i results = _mm512_setzero_si512();
    __m512i nc[8][4];
    for (int b = 0; b < m.nblocks; b++)
        for (int r = 0; r < 4; r++)
            nc[b][r] = _mm512_load_si512((const void *)m.comp[b][r]);
    for (size_t i = 0; i < n_rows; i++) {
        uint64_t q[4];
        std::memcpy(q, rows + 32 * i, 32);
        const __m512i q0 = _mm512_set1_epi64((long long)q[0]), q1 = _mm512_set1_epi64((long long)q[1]),
                      q2 = _mm512_set1_epi64((long long)q[2]), q3 = _mm512_set1_epi64((long long)q[3]);
        for (int b = 0; b < m.nblocks; b++) {
            __m512i c0 = _mm512_popcnt_epi64(_mm512_xor_si512(nc[b][0], q0));
            __m512i c1 = _mm512_popcnt_epi64(_mm512_xor_si512(nc[b][1], q1));
            __m512i c2 = _mm512_popcnt_epi64(_mm512_xor_si512(nc[b][2], q2));
            __m512i c3 = _mm512_popcnt_epi64(_mm512_xor_si512(nc[b][3], q3));
            __m512i sum = _mm512_add_epi64(_mm512_add_epi64(c0, c1), _mm512_add_epi64(c2, c3));
            if (FIND) {
                unsigned matched = _mm512_cmpge_epi64_mask(sum, thr);
                alignas(64) uint64_t sv[8];
                if (matched)
                    _mm512_store_si512((void *)sv, sum);
                while (matched) {
                    unsigned k = _tzcnt_u32(matched);
                    int nidx = b * 8 + (int)k;
                    if (nidx < n_valid)
                        hits->push_back({row_base + i, needle_base + (uint32_t)nidx,
                                         256u - (uint32_t)sv[k]});
                    matched ^= 1u << k;
                }
            } else {
                results = _mm512_or_si512(results, _mm512_add_epi64(sum, addend));
            }
        }
    }
    if (FIND)
        return !hits->empty();
    // scan: bit 8 set <=> 256 - dist + threshold >= 256 (matching/mod.rs:260-333)
    alignas(64) uint64_t o[8];
    _mm512_store_si512((void *)o, results);
    bool any = false;
    for (int k = 0; k < 8; k++)
        any |= (o[k] & 256) != 0;
    return any;
}

// scalar u64 path (matching/mod.rs:232-257, 336-362)
template <bool FIND>
bool scalar_run(const uint8_t *needles, int n_needles, uint64_t thr, const uint8_t *rows,
                size_t n_rows, size_t row_base, uint32_t needle_base, std::vector<Match> *hits)
{
    bool matched = false;
    for (size_t i = 0; i < n_rows; i++) {
        uint64_t q[4];
        std::memcpy(q, rows + 32 * i, 32);
        for (int j = 0; j < n_needles; j++) {
            uint64_t n[4];
            std::memcpy(n, needles + 32 * j, 32);
            uint64_t d = (uint64_t)__builtin_popcountll(q[0] ^ n[0]) + __builtin_popcountll(q[1] ^ n[1]) +
                         __builtin_popcountll(q[2] ^ n[2]) + __builtin_popcountll(q[3] ^ n[3]);
            if (d <= thr) {
                matched = true;
                if (FIND)
                    hits->push_back({row_base + i, needle_base + (uint32_t)j, (uint32_t)d});
            }
        }
    }
    return matched;
}

bool have_avx2()
{
    return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
}
bool have_avx512_popcnt()
{
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vpopcntdq") &&
           __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl");
}

} // namespace

/* bit 0: AVX2 hasher available, bit 1: AVX-512 VPOPCNTDQ matcher available */
API int cpu_baseline_isa(void)
{
    return (have_avx2() ? 1 : 0) | (have_avx512_popcnt() ? 2 : 0);
}

API int cpu_baseline_hw_threads(void)
{
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

/* Avx2F32Kernel::quantize (x86.rs:440-621) on one 16x16 PDQF: bisection threshold + 32 hash bytes.
 * Returns 0, or -1 when the host has no AVX2+FMA. */
API int cpu_baseline_quantize_avx2(const float *pdqf, float *threshold, uint8_t *out)
{
    if (!have_avx2())
        return -1;
    quantize_avx2(pdqf, threshold, out);
    return 0;
}

/* Hash n frames (8 dihedral hashes each) with the AVX2 kernel on `threads` threads.
 * hashes [n][8][32], quality [n][8], thresholds [n][8], pdqf [n][256] or NULL.
 * Returns 0, or -1 when the host has no AVX2+FMA. */
API int cpu_baseline_hash_batch(const float *frames, size_t n, uint8_t *hashes, float *quality,
                                float *thresholds, float *pdqf, int threads)
{
    if (!have_avx2())
        return -1;
    parallel_for(n, threads, [=](size_t lo, size_t hi, int) {
        Scratch *s = new Scratch();
        for (size_t f = lo; f < hi; f++)
            hash_all_avx2(frames + f * 512 * 512, *s, hashes + f * 256, quality + f * 8,
                          thresholds + f * 8, pdqf ? pdqf + f * 256 : nullptr);
        delete s;
    });
    return 0;
}

/* Match q needles against n_rows rows the way the reference drives CpuMatcher:
 * haystack in 2048-row batches (examples/end_to_end.rs:243-248), needles in
 * groups of 32 (the reference's advice, matching/mod.rs:108-110).  find != 0:
 * all hits are collected, sorted by (row, needle), the first `cap` stored.
 * Returns the number of hits (find) or 0/1 (scan). force_scalar selects the
 * u64 popcnt path. */
API size_t cpu_baseline_match(const uint8_t *rows, size_t n_rows, const uint8_t *needles,
                              size_t q, uint32_t threshold, int find, Match *out, size_t cap,
                              int threads, int force_scalar)
{
    const bool simd = have_avx512_popcnt() && !force_scalar;
    const size_t B = 2048;
    const size_t n_batches = (n_rows + B - 1) / B;
    std::vector<std::vector<Match>> per(threads > 0 ? threads : 1);
    std::atomic<int> any{0};
    parallel_for(n_batches, threads, [&](size_t lo, size_t hi, int tid) {
        std::vector<Match> &hits = per[tid];
        std::vector<Avx512Matcher> ms;
        if (simd) {
            for (size_t g = 0; g < q; g += 32) {
                Avx512Matcher m;
                avx512_build(m, needles + 32 * g, (int)std::min<size_t>(32, q - g), threshold);
                ms.push_back(m);
            }
        }
        bool matched = false;
        for (size_t b = lo; b < hi; b++) {
            const uint8_t *r = rows + 32 * B * b;
            size_t nr = std::min(B, n_rows - B * b);
            for (size_t g = 0, gi = 0; g < q; g += 32, gi++) {
                int nn = (int)std::min<size_t>(32, q - g);
                if (simd) {
                    if (find)
                        matched |= avx512_run<true>(ms[gi], nn, r, nr, B * b, (uint32_t)g, &hits);
                    else if (nn % 8 == 0)
                        matched |= avx512_run<false>(ms[gi], nn, r, nr, B * b, (uint32_t)g, nullptr);
                    else { // padded lanes would alias; use find for ragged groups
                        std::vector<Match> tmp;
                        matched |= avx512_run<true>(ms[gi], nn, r, nr, B * b, (uint32_t)g, &tmp);
                    }
                } else {
                    if (find)
                        matched |= scalar_run<true>(needles + 32 * g, nn, threshold, r, nr, B * b,
                                                    (uint32_t)g, &hits);
                    else
                        matched |= scalar_run<false>(needles + 32 * g, nn, threshold, r, nr, B * b,
                                                     (uint32_t)g, nullptr);
                }
            }
        }
        if (matched)
            any.store(1);
    });
    if (!find)
        return (size_t)any.load();
    std::vector<Match> all;
    for (auto &v : per)
        all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end(), [](const Match &a, const Match &b) {
        return a.row != b.row ? a.row < b.row : a.needle < b.needle;
    });
    for (size_t i = 0; i < all.size() && i < cap; i++)
        out[i] = all[i];
    return all.size();
}
