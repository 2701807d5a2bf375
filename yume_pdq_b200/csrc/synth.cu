// synth.cu -- device-side synthetic workload generators (benches and tests).
// Bit-identical restatements of oracle_synth_db / oracle_synth_pixel in
// oracle/pdq_oracle.c (integer-only, so CPU and GPU agree exactly); they follow the
// workload definitions of SURVEY.md 8(d).  Not part of the reference.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace ypdq {
namespace {

__device__ __forceinline__ uint64_t splitmix64_mix(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void synth_rows_kernel(uint64_t *rows, uint64_t n_words, uint64_t first_row, uint64_t seed)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n_words; i += stride)
        rows[i] = splitmix64_mix(seed + 4 * first_row + i);  // 4*(first+r) + w, i = 4r + w
}

__global__ void synth_frames_kernel(float *frames, uint64_t n_pixels, uint64_t first_frame,
                                    uint64_t seed, int kind)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n_pixels; i += stride) {
        const uint64_t frame = first_frame + (i >> 18);
        const uint32_t y = (uint32_t)(i >> 9) & 511u, x = (uint32_t)i & 511u;
        const uint64_t h = splitmix64_mix(seed + (frame << 18) + ((uint64_t)y << 9) + x);
        const uint32_t nz = (uint32_t)(h >> 56);
        float v;
        if (kind == 0) {
            v = (float)nz;
        } else {
            const uint64_t g = splitmix64_mix(seed ^ splitmix64_mix(frame + 0x1234u));
            uint32_t acc = 0;
#pragma unroll
            for (uint32_t t = 0; t < 4; t++) {
                const uint32_t a = (uint32_t)(g >> (12 * t)) & 7u;
                const uint32_t b = (uint32_t)(g >> (12 * t + 3)) & 7u;
                const uint32_t c = (uint32_t)(g >> (12 * t + 6)) & 63u;
                const uint32_t ph = (2u * a * x + 2u * b * y + 16u * c) & 1023u;
                const uint32_t tri = ph < 512u ? ph : 1024u - ph;
                acc += tri * (4u - t);
            }
            v = __fmul_rn((float)(3u * acc + nz), 1.0f / 64.0f);
        }
        frames[i] = v;
    }
}

// 8-bit frames (LUMA8 / RGB8 / RGBA8): channel c of pixel (y, x) = (3 * acc_c + noise_c) >> 6 with the
// triangle waves of kind 1 weighted 1 + ((t + c) & 3); alpha = a noise byte (the conversion ignores it).
// Restated by oracle_synth_pixels in oracle/pdq_oracle.c.
__global__ void synth_pixels_kernel(uint8_t *frames, uint64_t n_pixels, uint64_t first_frame, uint64_t seed,
                                    int channels)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n_pixels; i += stride) {
        const uint64_t frame = first_frame + (i >> 18);
        const uint32_t y = (uint32_t)(i >> 9) & 511u, x = (uint32_t)i & 511u;
        const uint64_t h = splitmix64_mix(seed + (frame << 18) + ((uint64_t)y << 9) + x);
        const uint64_t g = splitmix64_mix(seed ^ splitmix64_mix(frame + 0x1234u));
        uint32_t tri[4];
#pragma unroll
        for (uint32_t t = 0; t < 4; t++) {
            const uint32_t a = (uint32_t)(g >> (12 * t)) & 7u;
            const uint32_t b = (uint32_t)(g >> (12 * t + 3)) & 7u;
            const uint32_t c = (uint32_t)(g >> (12 * t + 6)) & 63u;
            const uint32_t ph = (2u * a * x + 2u * b * y + 16u * c) & 1023u;
            tri[t] = ph < 512u ? ph : 1024u - ph;
        }
        uint8_t *px = frames + i * channels;
        for (int c = 0; c < channels; c++) {
            if (c == 3) {
                px[c] = (uint8_t)(h >> 24);
                break;
            }
            uint32_t acc = 0;
#pragma unroll
            for (uint32_t t = 0; t < 4; t++)
                acc += tri[t] * (1u + ((t + (uint32_t)c) & 3u));
            px[c] = (uint8_t)((3u * acc + ((uint32_t)(h >> (56 - 8 * c)) & 0xffu)) >> 6);
        }
    }
}

}  // namespace

int launch_synth_pixels(uint8_t *frames, int pixel_format, size_t n, uint64_t first, uint64_t seed,
                        cudaStream_t s)
{
    if (n == 0)
        return 0;
    const int fmt = pixel_format & ~YUME_PDQ_CUDA_PIXEL_FMA;
    const int channels = fmt == YUME_PDQ_CUDA_PIXEL_LUMA8 ? 1 : fmt == YUME_PDQ_CUDA_PIXEL_RGB8 ? 3
                         : fmt == YUME_PDQ_CUDA_PIXEL_RGBA8 ? 4 : 0;
    if (!channels)
        return YUME_PDQ_CUDA_E_INVALID;
    const uint64_t px = (uint64_t)n << 18;
    unsigned grid = (unsigned)std::min<uint64_t>((px + 255) / 256, 148 * 32);
    synth_pixels_kernel<<<grid, 256, 0, s>>>(frames, px, first, seed, channels);
    YPDQ_LAUNCH_CHECK();
    return 0;
}

int launch_synth_rows(uint8_t *rows, size_t n, uint64_t first, uint64_t seed, cudaStream_t s)
{
    if (n == 0)
        return 0;
    const uint64_t words = (uint64_t)n * 4;
    unsigned grid = (unsigned)std::min<uint64_t>((words + 255) / 256, 148 * 32);
    synth_rows_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<uint64_t *>(rows), words, first, seed);
    YPDQ_LAUNCH_CHECK();
    return 0;
}

int launch_synth_frames(float *frames, size_t n, uint64_t first, uint64_t seed, int kind,
                        cudaStream_t s)
{
    if (n == 0)
        return 0;
    const uint64_t px = (uint64_t)n << 18;
    unsigned grid = (unsigned)std::min<uint64_t>((px + 255) / 256, 148 * 32);
    synth_frames_kernel<<<grid, 256, 0, s>>>(frames, px, first, seed, kind);
    YPDQ_LAUNCH_CHECK();
    return 0;
}

}  // namespace ypdq
