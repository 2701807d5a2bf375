/*
 * yume_pdq.h -- the reference crate's C ABI (eternal-flame-AD/yume-pdq, include/yume_pdq.h:22-150,
 * generated there by cbindgen from src/lib.rs:136-284), declared for the CUDA implementation.
 *
 * Same names, types, argument order and buffer contracts, so a C/C++/ctypes user of the reference
 * compiles against this header and links `-lyume_pdq` (yume_pdq_b200/libyume_pdq.so is a link to
 * libyume_pdq_cuda.so) without touching its source.  What differs is documented at each symbol.
 * The batched GPU entry points live in yume_pdq_cuda.h, which includes this file.
 */
#ifndef YUME_PDQ_H
#define YUME_PDQ_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference include/yume_pdq.h:22-25 (emitted by build.rs:389-429) */
extern const uint32_t YUME_PDQ_VERSION_MAJOR;
extern const uint32_t YUME_PDQ_VERSION_MINOR;
extern const uint32_t YUME_PDQ_VERSION_PATCH;
extern const char YUME_PDQ_VERSION_STRING[];

/* reference include/yume_pdq.h:46-56, src/lib.rs:85-134: a dihedral as the 2x2 matrix
 * [dx.0, dx.1, dy.0, dy.1] of signed bytes, packed big-endian into 32 bits */
#ifndef YUME_PDQ_DIHEDRALS_DEFINED /* lets a translation unit that already has the reference's definition include this file */
#define YUME_PDQ_DIHEDRALS_DEFINED
typedef struct YumePDQ_Dihedrals {
    uint32_t packed;
} YumePDQ_Dihedrals;
#endif

/* reference include/yume_pdq.h:62-100, src/lib.rs:154-177 */
extern const struct YumePDQ_Dihedrals YUME_PDQ_DIHEDRAL_NORMAL;
extern const struct YumePDQ_Dihedrals YUME_PDQ_DIHEDRAL_FLIPPED;
extern const struct YumePDQ_Dihedrals YUME_PDQ_DIHEDRAL_FLOPPED;
extern const struct YumePDQ_Dihedrals YUME_PDQ_DIHEDRAL_ROTATED_180;
extern const struct YumePDQ_Dihedrals YUME_PDQ_DIHEDRAL_ROTATED_90;
extern const struct YumePDQ_Dihedrals YUME_PDQ_DIHEDRAL_ROTATED_270;
extern const struct YumePDQ_Dihedrals YUME_PDQ_DIHEDRAL_ROTATED_90_FLOPPED;
extern const struct YumePDQ_Dihedrals YUME_PDQ_DIHEDRAL_FLOPPED_ROTATED_270;

/* reference include/yume_pdq.h:44, src/lib.rs:186-187.  `output` and `pdqf` of the visiting call are
 * only valid until the callback returns; return false to stop. */
typedef bool (*DihedralCallback)(void *ctx, uint32_t dihedral, float threshold, float quality);

/* reference include/yume_pdq.h:120-125, src/lib.rs:251-283 (hash_get_threshold, lib.rs:427-455).
 * input: 512*512 f32, row-major; threshold: nullable; output: 32 bytes; buf1: 128*128 f32 (receives
 * the 127x127 tent-filtered image, zero padded); tmp: 128 f32 (scratch in the reference, left
 * untouched here); pdqf: 256 f32.  Returns the quality in [0,1].
 * Here: runs on CUDA device 0 (or $YUME_PDQ_CUDA_DEVICE); one upload of the frame, one download of
 * the results per call.  Returns NaN when no sm_100 device is usable (the reference has no error
 * channel; there is no CPU fallback).  *threshold is the scalar kernel's lower median (Torben,
 * kernel/mod.rs:212-253); with YUME_PDQ_CUDA_THRESHOLD=avx2 in the environment it is the value the
 * reference's AVX2 / AVX-512 kernels report instead (x86.rs:440-621) -- same hash bits. */
float yume_pdq_hash_smart_kernel(const float *input, float *threshold, uint8_t *output, float *buf1,
                                 float *tmp, float *pdqf);

/* reference include/yume_pdq.h:143-147, src/lib.rs:205-232 (visit_dihedrals, lib.rs:358-420).
 * Calls `callback` 7 times in the order FLOPPED, ROTATED_180, FLIPPED, ROTATED_90_FLOPPED,
 * ROTATED_270, FLOPPED_ROTATED_270, ROTATED_90; during each call `output` holds that dihedral's
 * hash and `pdqf` its transformed PDQF.  The callback receives the ORIGINAL `threshold` argument
 * every time (lib.rs:223-224).  Returns false iff the callback returned false (or no device is
 * usable); pdqf is left in the state of the last visited dihedral (C.P^T after all seven). */
bool yume_pdq_visit_dihedrals_smart_kernel(void *ctx, float threshold, uint8_t *output, float *pdqf,
                                           DihedralCallback callback);

#ifdef __cplusplus
}
#endif

#endif /* YUME_PDQ_H */
