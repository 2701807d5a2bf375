/*
 * pdq_oracle.c -- CPU restatement of the yume-pdq hash-and-match hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA
 * kernels in yume_pdq_b200/csrc.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product
 * library (libyume_pdq_cuda.so) never links or calls it.
 *
 * Every function restates one reference function of /root/reference (the
 * Rust crate cannot be built here: no rustc/cargo in the image).  Arithmetic
 * is strict IEEE binary32 in the reference's own summation order, so build
 * with:  gcc -O2 -ffp-contract=off -fno-fast-math  (see oracle/Makefile).
 *
 * Parity pin (tests/test_oracle_kat.py): README doctest vector
 * (README.md:313-322 -> hash README.md:432-436, PDQF table README.md:397-412),
 * test-data/aaa-orig.jpg -> test-data/aaa-orig.jpg.expect, the re-threshold
 * identity of lib.rs:659-687 and the planted-positive matcher procedure of
 * matching/mod.rs:460-538.
 *
 * The look-up tables come from include/yume_pdq_luts.h (oracle/lut_gen.py).
 */
#include <math.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../include/yume_pdq_luts.h"

#define API __attribute__((visibility("default")))

static inline float lut_f32(uint32_t bits)
{
    float f;
    memcpy(&f, &bits, sizeof f);
    return f;
}

/* ------------------------------------------------------------------ a1 --
 * DefaultKernel::jarosz_compress, src/kernel/mod.rs:644-666.
 * out is [127][out_stride]; only the 127x127 region is written. */
API void oracle_jarosz_compress(const float *in, float *out, size_t out_stride)
{
    float w[49];
    for (int i = 0; i < 49; i++)
        w[i] = lut_f32(YUME_PDQ_TENT_WEIGHT_BITS[i]);
    for (int outi = 0; outi < 127; outi++) {
        size_t in_i = (size_t)YUME_PDQ_CONV_OFFSET[outi] - YUME_PDQ_TENT_OFFSET;
        for (int outj = 0; outj < 127; outj++) {
            size_t in_j = (size_t)YUME_PDQ_CONV_OFFSET[outj] - YUME_PDQ_TENT_OFFSET;
            float sum = 0.0f;
            for (int di = 0; di < YUME_PDQ_TENT_ROWS; di++)
                for (int dj = 0; dj < YUME_PDQ_TENT_COLS; dj++)
                    sum += w[di * YUME_PDQ_TENT_COLS + dj] * in[(in_i + di) * 512 + in_j + dj];
            out[(size_t)outi * out_stride + outj] = sum;
        }
    }
}

/* ------------------------------------------------------------------ a2 --
 * DefaultKernel::dct2d_impl, src/kernel/mod.rs:543-574: 4 round-robin partial
 * sums per dot product, combined ((s0+s1)+s2)+s3. tmp has >= 127 floats. */
API void oracle_dct2d(const float *buf, size_t stride, float *tmp, float *pdqf)
{
    static float d[16 * 127];
    static int init = 0;
    if (!init) {
        for (int i = 0; i < 16 * 127; i++)
            d[i] = lut_f32(YUME_PDQ_DCT_MATRIX_BITS[i]);
        init = 1;
    }
    for (int k = 0; k < 16; k++) {
        for (int j = 0; j < 127; j++) {
            float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (int k2 = 0; k2 < 127; k2++)
                s[k2 & 3] += d[k * 127 + k2] * buf[(size_t)k2 * stride + j];
            tmp[j] = s[0] + s[1] + s[2] + s[3];
        }
        for (int j = 0; j < 16; j++) {
            float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (int m = 0; m < 127; m++)
                s[m & 3] += tmp[m] * d[j * 127 + m];
            pdqf[k * 16 + j] = s[0] + s[1] + s[2] + s[3];
        }
    }
}

/* ------------------------------------------------------------------ a3 --
 * sum_of_gradients, src/kernel/mod.rs:617-642; adjust_quality 611-615. */
API float oracle_sum_of_gradients(const float *p)
{
    float g = 0.0f;
    for (int i = 0; i < 15; i++)
        for (int j = 0; j < 16; j++)
            g += fabsf(p[i * 16 + j] - p[(i + 1) * 16 + j]);
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 15; j++)
            g += fabsf(p[i * 16 + j] - p[i * 16 + j + 1]);
    return g;
}

API float oracle_adjust_quality(float g)
{
    float scaled = g / 180.0f; /* QUALITY_ADJUST_DIVISOR, mod.rs:257 */
    return scaled < 1.0f ? scaled : (scaled != scaled ? 1.0f : 1.0f); /* f32::min(NaN,1)=1 */
}

/* ------------------------------------------------------------------ a4 --
 * torben_median, src/kernel/mod.rs:212-253 (half = (256+1)/2 = 128). */
API float oracle_torben_median(const float *m)
{
    float min = m[0], max = m[0];
    for (int i = 1; i < 256; i++) {
        min = fminf(min, m[i]);
        max = fmaxf(max, m[i]);
    }
    const int half = (16 * 16 + 1) / 2;
    for (;;) {
        float guess = (min + max) / 2.0f;
        int less = 0, greater = 0, equal = 0;
        float maxltguess = min, mingtguess = max;
        for (int i = 0; i < 256; i++) {
            float v = m[i];
            if (v < guess) {
                less++;
                if (v > maxltguess)
                    maxltguess = v;
            } else if (v > guess) {
                greater++;
                if (v < mingtguess)
                    mingtguess = v;
            } else {
                equal++;
            }
        }
        if (less <= half && greater <= half) {
            if (less >= half)
                return maxltguess;
            else if (less + equal >= half)
                return guess;
            else
                return mingtguess;
        } else if (less > greater) {
            max = maxltguess;
        } else {
            min = mingtguess;
        }
    }
}

/* DefaultKernel::quantize, src/kernel/mod.rs:595-608: bit i (flat 16*row+col)
 * -> out[31 - i/8] bit (i % 8), set iff value > median (strict). */
API void oracle_quantize(const float *p, float *threshold, uint8_t *out)
{
    float median = oracle_torben_median(p);
    *threshold = median;
    memset(out, 0, 32);
    for (int i = 0; i < 256; i++)
        if (p[i] > median)
            out[32 - 1 - i / 8] += (uint8_t)(1u << (i % 8));
}

/* ------------------------------------------------------------------ a5 --
 * threshold_2d_f32, src/kernel/threshold.rs:37-119 (same packing, strict >). */
API void oracle_threshold_2d_f32(const float *p, uint8_t *out, float threshold)
{
    for (int byte = 0; byte < 32; byte++) {
        uint8_t w = 0;
        for (int b = 0; b < 8; b++)
            if (p[byte * 8 + b] > threshold)
                w |= (uint8_t)(1u << b);
        out[31 - byte] = w;
    }
}

/* ------------------------------------------------------------- a6..a8 --
 * pdqf_negate_alt_cols::<NEGATE>, src/kernel/mod.rs:368-382 (AVX2 330-376
 * agrees): columns j = (NEGATE ? 1 : 0), +2, ... */
API void oracle_negate_alt_cols(float *p, int negate)
{
    for (int i = 0; i < 16; i++)
        for (int j = negate ? 1 : 0; j < 16; j += 2)
            p[i * 16 + j] = -p[i * 16 + j];
}

/* pdqf_negate_alt_rows::<NEGATE>, AVX2/AVX-512 convention
 * (src/kernel/x86.rs:285-326, 989-1024): NEGATE=true starts at row 0.  This is
 * what smart_kernel() runs on x86 and the contract chosen in SURVEY.md 8(a7). */
API void oracle_negate_alt_rows(float *p, int negate)
{
    for (int i = negate ? 0 : 1; i < 16; i += 2)
        for (int j = 0; j < 16; j++)
            p[i * 16 + j] = -p[i * 16 + j];
}

/* Trait-default variant (src/kernel/mod.rs:391-405): NEGATE=true starts at row
 * 1.  Kept only so tests can document the disagreement; not on the contract. */
API void oracle_negate_alt_rows_trait_default(float *p, int negate)
{
    for (int i = negate ? 1 : 0; i < 16; i += 2)
        for (int j = 0; j < 16; j++)
            p[i * 16 + j] = -p[i * 16 + j];
}

/* pdqf_t, src/kernel/mod.rs:345-361 */
API void oracle_pdqf_t(float *p)
{
    for (int i = 0; i < 16; i++)
        for (int j = i + 1; j < 16; j++) {
            float t = p[i * 16 + j];
            p[i * 16 + j] = p[j * 16 + i];
            p[j * 16 + i] = t;
        }
}

/* pdqf_negate_off_diagonals, src/kernel/mod.rs:413-427: (j - i) wrapping % 2 == 1 */
API void oracle_negate_off_diagonals(float *p)
{
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 16; j++)
            if (((size_t)j - (size_t)i) % 2 == 1)
                p[i * 16 + j] = -p[i * 16 + j];
}

/* ----------------------------------------------------------------- a11 --
 * hash_get_threshold, src/lib.rs:427-455.  buf1 is [128][128] (zeroed padding),
 * tmp is [128]. threshold may be NULL (ffi, lib.rs:272-277). */
API float oracle_hash_get_threshold(const float *in, float *threshold, uint8_t *out,
                                    float *buf1, float *tmp, float *pdqf)
{
    float dummy;
    oracle_jarosz_compress(in, buf1, 128);
    oracle_dct2d(buf1, 128, tmp, pdqf);
    float quality = oracle_adjust_quality(oracle_sum_of_gradients(pdqf));
    oracle_quantize(pdqf, threshold ? threshold : &dummy, out);
    return quality;
}

/* ------------------------------------------------------------ a9, a10 --
 * Dihedrals::from_tuples, src/lib.rs:85-134 */
#define DIH(dx0, dx1, dy0, dy1)                                                      \
    ((uint32_t)(uint8_t)(dx0) << 24 | (uint32_t)(uint8_t)(dx1) << 16 |               \
     (uint32_t)(uint8_t)(dy0) << 8 | (uint32_t)(uint8_t)(dy1))
API const uint32_t ORACLE_DIHEDRAL_ORDER[8] = {
    DIH(1, 0, 0, 1),   /* NORMAL */
    DIH(-1, 0, 0, 1),  /* FLOPPED */
    DIH(-1, 0, 0, -1), /* ROTATED_180 */
    DIH(1, 0, 0, -1),  /* FLIPPED */
    DIH(0, -1, -1, 0), /* ROTATED_90_FLOPPED */
    DIH(0, 1, 1, 0),   /* ROTATED_270 */
    DIH(0, -1, 1, 0),  /* FLOPPED_ROTATED_270 */
    DIH(0, 1, -1, 0),  /* ROTATED_90 */
};

typedef bool (*oracle_dihedral_cb)(void *ctx, uint32_t dihedral, float threshold, float quality);

/* visit_dihedrals, src/lib.rs:358-420.  `cb` receives the per-dihedral
 * threshold (the Rust closure signature); the C ABI wrapper in lib.rs:205-232
 * forwards the ORIGINAL threshold instead -- tests cover both. */
API bool oracle_visit_dihedrals(void *ctx, float threshold, uint8_t *out, float *pdqf,
                                oracle_dihedral_cb cb)
{
#define CALLBACK(idx, thr)                                                           \
    do {                                                                             \
        float q = oracle_adjust_quality(oracle_sum_of_gradients(pdqf));              \
        if (!cb(ctx, ORACLE_DIHEDRAL_ORDER[idx], (thr), q))                          \
            return false;                                                            \
    } while (0)
    float thr_cols = threshold, thr_rows = threshold, thr_od = threshold;
    oracle_negate_alt_cols(pdqf, 0);
    oracle_quantize(pdqf, &thr_cols, out);
    CALLBACK(1, thr_cols);
    oracle_negate_alt_rows(pdqf, 1);
    oracle_quantize(pdqf, &thr_od, out);
    CALLBACK(2, thr_od);
    oracle_negate_alt_cols(pdqf, 0);
    oracle_quantize(pdqf, &thr_rows, out);
    CALLBACK(3, thr_rows);
    oracle_negate_alt_rows(pdqf, 1);
    oracle_pdqf_t(pdqf);
    oracle_threshold_2d_f32(pdqf, out, threshold);
    CALLBACK(4, threshold);
    oracle_negate_alt_rows(pdqf, 1);
    oracle_threshold_2d_f32(pdqf, out, thr_cols);
    CALLBACK(5, thr_cols);
    oracle_negate_alt_cols(pdqf, 0);
    oracle_threshold_2d_f32(pdqf, out, thr_od);
    CALLBACK(6, thr_od);
    oracle_negate_alt_rows(pdqf, 1);
    oracle_threshold_2d_f32(pdqf, out, thr_rows);
    CALLBACK(7, thr_rows);
    return true;
#undef CALLBACK
}

struct collect_ctx {
    uint8_t *hashes;     /* [8][32] */
    float *quality;      /* [8] */
    float *thresholds;   /* [8] per-dihedral threshold as the Rust closure sees it */
    const uint8_t *out;
    int i;
};

static bool collect_cb(void *ctx_, uint32_t dihedral, float threshold, float quality)
{
    struct collect_ctx *c = (struct collect_ctx *)ctx_;
    (void)dihedral;
    memcpy(c->hashes + 32 * c->i, c->out, 32);
    c->quality[c->i] = quality;
    c->thresholds[c->i] = threshold;
    c->i++;
    return true;
}

/* Convenience driver used by the parity tests: one frame -> 8 hashes in the
 * order of ORACLE_DIHEDRAL_ORDER, 8 qualities, 8 thresholds, and the original
 * PDQF (pdqf_out, 256 floats; may be NULL). */
API void oracle_hash_all_dihedrals(const float *in, uint8_t *hashes, float *quality,
                                   float *thresholds, float *pdqf_out)
{
    static __thread float buf1[128 * 128];
    float tmp[128];
    float pdqf[256];
    uint8_t out[32];
    memset(buf1, 0, sizeof buf1);
    memset(tmp, 0, sizeof tmp);
    float thr = 0.0f;
    quality[0] = oracle_hash_get_threshold(in, &thr, out, buf1, tmp, pdqf);
    thresholds[0] = thr;
    memcpy(hashes, out, 32);
    if (pdqf_out)
        memcpy(pdqf_out, pdqf, sizeof pdqf);
    struct collect_ctx c = {hashes, quality, thresholds, out, 1};
    oracle_visit_dihedrals(&c, thr, out, pdqf, collect_cb);
}

/* ------------------------------------------------------------ a13, a15 --
 * CpuMatcher::find (scalar), src/matching/mod.rs:336-362: rows outer, needles
 * inner, hit iff sum of 4 popcnt64 <= threshold (inclusive).  Reports every
 * hit in (row, needle) lexicographic order; returns the total number of hits
 * (the first `cap` are stored). */
struct oracle_match {
    uint64_t row;
    uint32_t needle;
    uint32_t dist;
};

API size_t oracle_matcher_find(const uint8_t *rows, size_t n_rows, const uint8_t *needles,
                               size_t n_needles, uint32_t threshold, struct oracle_match *out,
                               size_t cap)
{
    size_t found = 0;
    for (size_t i = 0; i < n_rows; i++) {
        uint64_t r[4];
        memcpy(r, rows + 32 * i, 32);
        for (size_t j = 0; j < n_needles; j++) {
            uint64_t n[4];
            memcpy(n, needles + 32 * j, 32);
            uint64_t dist = 0;
            for (int w = 0; w < 4; w++)
                dist += (uint64_t)__builtin_popcountll(r[w] ^ n[w]);
            if (dist <= (uint64_t)threshold) {
                if (found < cap) {
                    out[found].row = i;
                    out[found].needle = (uint32_t)j;
                    out[found].dist = (uint32_t)dist;
                }
                found++;
            }
        }
    }
    return found;
}

/* CpuMatcher::scan (scalar), src/matching/mod.rs:232-257 */
API int oracle_matcher_scan(const uint8_t *rows, size_t n_rows, const uint8_t *needles,
                            size_t n_needles, uint32_t threshold)
{
    return oracle_matcher_find(rows, n_rows, needles, n_needles, threshold, NULL, 0) != 0;
}

/* VulkanMatcher::find, src/matching/vulkan.rs:408-472: atomicMin of
 * needle + row * n_needles over all hits; returns UINT64_MAX when none. */
API uint64_t oracle_matcher_find_min(const uint8_t *rows, size_t n_rows, const uint8_t *needles,
                                     size_t n_needles, uint32_t threshold)
{
    struct oracle_match m;
    size_t n = oracle_matcher_find(rows, n_rows, needles, n_needles, threshold, &m, 1);
    return n ? m.row * n_needles + m.needle : UINT64_MAX;
}

/* ------------------------------------------------- synthetic workloads --
 * Counter-based generators shared (by restatement) with the device-side
 * generators in yume_pdq_b200/csrc/synth.cu, SURVEY.md 8(d):
 * DB row r, 64-bit word w = splitmix64_mix(seed + 4r + w). */
static inline uint64_t splitmix64_mix(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

API void oracle_synth_db(uint8_t *rows, size_t n_rows, uint64_t first_row, uint64_t seed)
{
    for (size_t i = 0; i < n_rows; i++) {
        uint64_t w[4];
        for (int k = 0; k < 4; k++)
            w[k] = splitmix64_mix(seed + 4 * (first_row + i) + (uint64_t)k);
        memcpy(rows + 32 * i, w, 32);
    }
}

/* Frame f, pixel (y, x): integer-only generators, so the CPU and the device
 * produce bit-identical f32 pixels.
 *   kind 0: uniform 8-bit noise, value = top byte of mix(seed + f*2^18 + y*512 + x)
 *   kind 1: smooth: four integer triangle waves a*x + b*y (0..7 cycles per
 *           frame each, per-frame random) plus noise, value = (3*acc + n) / 64
 *           in [0, 244] -- low-frequency content so the 16x16 PDQF is well
 *           conditioned (uniform noise alone gives a tiny median gap). */
API float oracle_synth_pixel(uint64_t seed, uint64_t frame, uint32_t y, uint32_t x, int kind)
{
    uint64_t h = splitmix64_mix(seed + (frame << 18) + ((uint64_t)y << 9) + x);
    uint32_t n = (uint32_t)(h >> 56);
    if (kind == 0)
        return (float)n;
    uint64_t g = splitmix64_mix(seed ^ splitmix64_mix(frame + 0x1234u));
    uint32_t acc = 0;
    for (uint32_t t = 0; t < 4; t++) {
        uint32_t a = (uint32_t)(g >> (12 * t)) & 7u;
        uint32_t b = (uint32_t)(g >> (12 * t + 3)) & 7u;
        uint32_t c = (uint32_t)(g >> (12 * t + 6)) & 63u;
        uint32_t ph = (2u * a * x + 2u * b * y + 16u * c) & 1023u;
        uint32_t tri = ph < 512u ? ph : 1024u - ph;
        acc += tri * (4u - t);
    }
    return (float)(3u * acc + n) * (1.0f / 64.0f);
}

API void oracle_synth_frames(float *frames, size_t n_frames, uint64_t first_frame, uint64_t seed,
                             int kind)
{
    for (size_t f = 0; f < n_frames; f++)
        for (uint32_t y = 0; y < 512; y++)
            for (uint32_t x = 0; x < 512; x++)
                frames[(f * 512 + y) * 512 + x] =
                    oracle_synth_pixel(seed, first_frame + f, y, x, kind);
}

/* 8-bit frames for the pixel-format paths (restated by synth_pixels_kernel in synth.cu): channel c of pixel
 * (y, x) = (3 * acc_c + noise_c) >> 6 with the triangle waves of kind 1 weighted 1 + ((t + c) & 3);
 * alpha = a noise byte.  channels = 1 (LUMA8), 3 (RGB8) or 4 (RGBA8). */
API void oracle_synth_pixels(uint8_t *frames, size_t n_frames, uint64_t first_frame, uint64_t seed,
                             int channels)
{
    for (size_t f = 0; f < n_frames; f++) {
        uint64_t frame = first_frame + f;
        uint64_t g = splitmix64_mix(seed ^ splitmix64_mix(frame + 0x1234u));
        for (uint32_t y = 0; y < 512; y++)
            for (uint32_t x = 0; x < 512; x++) {
                uint64_t h = splitmix64_mix(seed + (frame << 18) + ((uint64_t)y << 9) + x);
                uint32_t tri[4];
                for (uint32_t t = 0; t < 4; t++) {
                    uint32_t a = (uint32_t)(g >> (12 * t)) & 7u;
                    uint32_t b = (uint32_t)(g >> (12 * t + 3)) & 7u;
                    uint32_t c = (uint32_t)(g >> (12 * t + 6)) & 63u;
                    uint32_t ph = (2u * a * x + 2u * b * y + 16u * c) & 1023u;
                    tri[t] = ph < 512u ? ph : 1024u - ph;
                }
                uint8_t *px = frames + ((f * 512 + y) * 512 + x) * (size_t)channels;
                for (int c = 0; c < channels; c++) {
                    if (c == 3) {
                        px[c] = (uint8_t)(h >> 24);
                        break;
                    }
                    uint32_t acc = 0;
                    for (uint32_t t = 0; t < 4; t++)
                        acc += tri[t] * (1u + ((t + (uint32_t)c) & 3u));
                    px[c] = (uint8_t)((3u * acc + ((uint32_t)(h >> (56 - 8 * c)) & 0xffu)) >> 6);
                }
            }
    }
}

/* ------------------------------------------------------------------ f2 --
 * Input conversion, the step before hash_get_threshold in the reference CLI (main.rs:486-608).
 * LUMA8: `row_buf[j] as f32` (main.rs:508-510).
 * Kernel::cvt_rgb8_to_luma8f / cvt_rgba8_to_luma8f, trait default (src/kernel/mod.rs:303-340):
 *   luma = cr * r + cg * g + cb * b, separate f32 multiplies and adds, left to right; alpha ignored.
 * fma != 0 restates the AVX2 override of the RGBA8 conversion (src/kernel/x86.rs:686-753):
 *   v = r * cr;  v = fma(g, cg, v);  v = fma(b, cb, v).
 * Coefficients are passed as f32 bit patterns like the trait's const generics; the CLI always uses
 * RGB8_TO_LUMA8_TABLE_ITU = [0.299, 0.587, 0.114] (src/kernel/constants.rs:23). */
API const float ORACLE_RGB8_TO_LUMA8_TABLE_ITU[3] = {0.299f, 0.587f, 0.114f};

API void oracle_cvt_to_luma8f(const uint8_t *in, int channels, size_t n_pixels, uint32_t r_bits,
                              uint32_t g_bits, uint32_t b_bits, int fma, float *out)
{
    const float cr = lut_f32(r_bits), cg = lut_f32(g_bits), cb = lut_f32(b_bits);
    for (size_t i = 0; i < n_pixels; i++) {
        const uint8_t *px = in + i * (size_t)channels;
        if (channels == 1) {
            out[i] = (float)px[0];
            continue;
        }
        float r = (float)px[0], g = (float)px[1], b = (float)px[2];
        if (fma) {
            float v = r * cr;
            v = fmaf(g, cg, v);
            v = fmaf(b, cb, v);
            out[i] = v;
        } else {
            float t0 = cr * r, t1 = cg * g, t2 = cb * b;
            out[i] = (t0 + t1) + t2;
        }
    }
}
