/*
 * yume_pdq_cuda.h -- C ABI of the B200-native (sm_100a) PDQ hash-and-match engine.
 *
 * Drop-in boundary for the hot path of eternal-flame-AD/yume-pdq.  Two groups
 * of entry points:
 *
 *  (A) the reference's own C ABI, same names / signatures / semantics
 *      (/root/reference include/yume_pdq.h:22-150, src/lib.rs:136-284), now
 *      executed on the GPU;
 *  (B) batched `yume_pdq_cuda_*` entry points that the Rust side
 *      (`impl Kernel for CudaKernel`, `impl PDQMatcher for CudaMatcher`,
 *      integration/rust/) and the Python/C++ hosts bind.  Per-frame trait calls
 *      cannot feed a GPU, so every stage of the `Kernel` trait
 *      (src/kernel/mod.rs:268-497) is exported in batched form.
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every function in
 * (B) returns an int: 0 = OK, > 0 = a cudaError_t value, < 0 = YUME_PDQ_CUDA_E_*.
 * Unless stated otherwise a data pointer may be a HOST pointer (pageable or
 * pinned) or a DEVICE pointer on `device`; the library detects which.  When all
 * data pointers are device pointers the call is asynchronous on `stream`
 * (a cudaStream_t passed as void*, NULL = default stream); with any host
 * pointer the call stages through the device and returns after completion.
 * There is no CPU fallback: without a usable CUDA device every call fails with
 * YUME_PDQ_CUDA_E_NO_DEVICE.
 */
#ifndef YUME_PDQ_CUDA_H
#define YUME_PDQ_CUDA_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

/* ------------------------------------------------------------------ (A) --
 * The reference's own ABI: version symbols, YumePDQ_Dihedrals and its 8 constants, DihedralCallback,
 * yume_pdq_hash_smart_kernel, yume_pdq_visit_dihedrals_smart_kernel. */
#include "yume_pdq.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ (B) -- */
#define YUME_PDQ_CUDA_OK 0
#define YUME_PDQ_CUDA_E_INVALID (-1)   /* bad argument (NULL, size, q == 0, ...) */
#define YUME_PDQ_CUDA_E_NO_DEVICE (-2) /* no usable CUDA device / not an sm_100 part */
#define YUME_PDQ_CUDA_E_CAPACITY (-3)  /* more matches than `cap`; *n_found holds the total (a lower bound once the
                                        * 32-bit device counter saturates at 0xF0000000 matches) */
#define YUME_PDQ_CUDA_E_NOMEM (-4)
#define YUME_PDQ_CUDA_E_TIMEOUT (-5)   /* a peer did not answer a mailbox query within 2 s */

const char *yume_pdq_cuda_error_string(int code);
int yume_pdq_cuda_device_count(void);
/* Kernel::ident (src/kernel/mod.rs:429-430) / PDQMatcher::ident (matching/mod.rs:63-64) */
const char *yume_pdq_cuda_kernel_ident(void);
const char *yume_pdq_cuda_matcher_ident(void);
/* number of kernel launches issued by this library since load (for bench.py's gpu_launches) */
uint64_t yume_pdq_cuda_launch_count(void);

/* Order of the 8 dihedral hashes produced per frame:
 * NORMAL, FLOPPED, ROTATED_180, FLIPPED, ROTATED_90_FLOPPED, ROTATED_270,
 * FLOPPED_ROTATED_270, ROTATED_90 (NORMAL + the visit order of lib.rs:358-420). */
extern const uint32_t YUME_PDQ_CUDA_DIHEDRAL_ORDER[8];

/* Page-locked host memory for callers without a CUDA runtime of their own (the `pipe` front end reads frames
 * straight into it): host-pointer calls on such buffers copy at the full PCIe rate and overlap with kernels. */
int yume_pdq_cuda_host_alloc(int device, size_t bytes, void **out);
int yume_pdq_cuda_host_free(void *p);

/* ---- hasher: fused hot path (hash_get_threshold + visit_dihedrals, batched) ----
 * frames   n x 512 x 512 f32
 * hashes   n x 8 x 32 bytes, dihedral order above          (required)
 * quality  n x 8 f32  (quality of each transformed PDQF)    (nullable)
 * thresholds n x 8 f32: the threshold each dihedral was quantised with
 *          (thr0, thr_c, thr_od, thr_r, thr0, thr_c, thr_od, thr_r)  (nullable)
 * pdqf     n x 256 f32, the untransformed PDQF             (nullable) */
int yume_pdq_cuda_hash_batch(int device, const float *frames, size_t n, uint8_t *hashes,
                             float *quality, float *thresholds, float *pdqf, void *stream);

/* ---- pixel formats: the input conversion fused into the hasher ----
 * The reference CLI converts every input row to f32 luma before hash_get_threshold
 * (src/main.rs:486-608): LUMA8 by `as f32` (main.rs:508-510), RGB8 / RGBA8 through
 * Kernel::cvt_rgb8_to_luma8f / cvt_rgba8_to_luma8f (src/kernel/mod.rs:303-340) with the ITU
 * BT.601 table (src/kernel/constants.rs:23).  Here the conversion happens inside the window
 * load of the hash kernel, so an 8-bit frame crosses PCIe and HBM once, as bytes.
 * A frame is 512 x 512 pixels, row-major, tightly packed (RGB8: 3 bytes per pixel R,G,B;
 * RGBA8: R,G,B,A with A ignored).  Results equal cvt-to-f32 followed by the f32 path, bit for bit.
 * YUME_PDQ_CUDA_PIXEL_FMA (RGB8/RGBA8 only) selects the rounding of the AVX2 override
 * (mul, fmadd, fmadd; src/kernel/x86.rs:686-753) instead of the trait default. */
#define YUME_PDQ_CUDA_PIXEL_F32 0
#define YUME_PDQ_CUDA_PIXEL_LUMA8 1
#define YUME_PDQ_CUDA_PIXEL_RGB8 2
#define YUME_PDQ_CUDA_PIXEL_RGBA8 3
#define YUME_PDQ_CUDA_PIXEL_FMA 0x100
/* bytes of one frame in pixel_format (0 = invalid format) */
size_t yume_pdq_cuda_frame_bytes(int pixel_format);
/* yume_pdq_cuda_hash_batch for frames in any pixel format (device pointers 16-byte aligned) */
int yume_pdq_cuda_hash_batch_pixels(int device, const void *frames, int pixel_format, size_t n,
                                    uint8_t *hashes, float *quality, float *thresholds, float *pdqf,
                                    void *stream);
/* Kernel::cvt_rgb8_to_luma8f / cvt_rgba8_to_luma8f (mod.rs:303-340) and the LUMA8 cast on their
 * own, batched over n_pixels pixels (any number of rows): pixels -> out[n_pixels] f32.
 * coeff = {R, G, B} like the trait's const generics, NULL = the ITU table. */
int yume_pdq_cuda_cvt_to_luma8f(int device, const void *pixels, int pixel_format, size_t n_pixels,
                                const float *coeff, float *out, void *stream);

/* ---- hasher: the stages of the `Kernel` trait, batched over n ----
 * Kernel::jarosz_compress (mod.rs:644-666): frames n x 512 x 512 -> buf1 n x 128 x 128
 * (127 x 127 valid, padding written as zero, cf. x86.rs:756-757). */
int yume_pdq_cuda_jarosz_compress(int device, const float *frames, size_t n, float *buf1,
                                  void *stream);
/* Kernel::dct2d (mod.rs:543-574): buf1 n x 128 x 128 -> pdqf n x 256 */
int yume_pdq_cuda_dct2d(int device, const float *buf1, size_t n, float *pdqf, void *stream);
/* Kernel::sum_of_gradients (mod.rs:617-642) and adjust_quality (mod.rs:611-615);
 * either output may be NULL. */
int yume_pdq_cuda_sum_of_gradients(int device, const float *pdqf, size_t n, float *gradient,
                                   float *quality, void *stream);
/* Kernel::quantize (mod.rs:595-608, Torben lower median): pdqf n x 256 ->
 * thresholds n, hashes n x 32 */
int yume_pdq_cuda_quantize(int device, const float *pdqf, size_t n, float *thresholds,
                           uint8_t *hashes, void *stream);
/* kernel::threshold::threshold_2d_f32 (threshold.rs:37-119) */
int yume_pdq_cuda_threshold_2d_f32(int device, const float *pdqf, size_t n,
                                   const float *thresholds, uint8_t *hashes, void *stream);
/* in-place PDQF transforms (mod.rs:345-427; row convention of x86.rs:285-326) */
#define YUME_PDQ_CUDA_OP_T 0                        /* pdqf_t */
#define YUME_PDQ_CUDA_OP_NEGATE_ALT_COLS_FALSE 1    /* pdqf_negate_alt_cols::<false>: cols 0,2,.. */
#define YUME_PDQ_CUDA_OP_NEGATE_ALT_COLS_TRUE 2     /* ::<true>: cols 1,3,.. */
#define YUME_PDQ_CUDA_OP_NEGATE_ALT_ROWS_TRUE 3     /* pdqf_negate_alt_rows::<true>: rows 0,2,.. */
#define YUME_PDQ_CUDA_OP_NEGATE_ALT_ROWS_FALSE 4    /* ::<false>: rows 1,3,.. */
#define YUME_PDQ_CUDA_OP_NEGATE_OFF_DIAGONALS 5     /* pdqf_negate_off_diagonals */
int yume_pdq_cuda_pdqf_transform(int device, float *pdqf, size_t n, int op, void *stream);
/* visit_dihedrals (lib.rs:358-420) batched: from pdqf n x 256 and the NORMAL
 * thresholds thr0[n] produce all 8 hashes / qualities / thresholds as hash_batch does. */
int yume_pdq_cuda_dihedrals_batch(int device, const float *pdqf, const float *thr0, size_t n,
                                  uint8_t *hashes, float *quality, float *thresholds,
                                  void *stream);

/* ---- matcher ------------------------------------------------------------
 * The haystack lives on the device (like VulkanVectorDatabase, vulkan.rs:40-65);
 * a query is a set of needles (like VulkanMatcher::scan/find, vulkan.rs:349-472),
 * but every match is reported (FindsAllMatches = B1 like CpuMatcher,
 * matching/mod.rs:222). */
typedef struct YumePDQ_CudaDb YumePDQ_CudaDb;

/* (row, needle, dist): row is the GLOBAL row index (local index + global_row_offset) */
typedef struct YumePDQ_Match {
    uint64_t row;
    uint32_t needle;
    uint32_t dist;
} YumePDQ_Match;

/* Upload n_rows x 32 bytes (host or device pointer).  global_row_offset is added
 * to every reported row so that a row-range shard per GPU reports global rows. */
int yume_pdq_cuda_db_create(int device, const uint8_t *rows, size_t n_rows,
                            uint64_t global_row_offset, YumePDQ_CudaDb **out);
/* Same, but the rows are generated on the device by the counter-based generator
 * (row r, 64-bit word w = splitmix64(seed + 4r + w), r = global index) --
 * benches use this for databases that cannot cross PCIe. */
int yume_pdq_cuda_db_create_synthetic(int device, size_t n_rows, uint64_t global_row_offset,
                                      uint64_t seed, YumePDQ_CudaDb **out);
int yume_pdq_cuda_db_destroy(YumePDQ_CudaDb *db);
size_t yume_pdq_cuda_db_len(const YumePDQ_CudaDb *db);
int yume_pdq_cuda_db_device(const YumePDQ_CudaDb *db);
/* Device pointers to the two planes the haystack is stored as: prefix[r] = bytes 0..15 of row r, suffix[r] = bytes
 * 16..31 (n_rows x 16 B each).  Prefix filters stream the first plane only. */
int yume_pdq_cuda_db_planes(const YumePDQ_CudaDb *db, const uint8_t **prefix, const uint8_t **suffix);
/* ALGO_AUTO's data-dependent state for this haystack (large needle batches at threshold <= 36): *fallback = 1 while
 * launches run the unfiltered tensor kernel because the prefix filter met too many candidates (near-duplicate
 * families); *probe_interval = launches between probes of the filter on a slice of the rows.  Diagnostic only. */
int yume_pdq_cuda_db_auto_state(const YumePDQ_CudaDb *db, int *fallback, uint32_t *probe_interval);
/* overwrite rows [first, first+n) (host or device source) -- used by tests to plant */
int yume_pdq_cuda_db_write_rows(YumePDQ_CudaDb *db, size_t first_local_row, const uint8_t *rows,
                                size_t n);

/* diagnostic: host-side microsecond stamps of the calling thread's last yume_pdq_hash_smart_kernel call (frame handed to
 * the copy engine, kernel launched, answer seen, outputs copied out) and, if kernel_clocks7 is not NULL, SM clocks since
 * its start at which the kernel's finishing CTA was elected, had the tent buffer in shared memory, finished dct stage 1,
 * stage 2, the median, the hash bits, the gradient sum */
int yume_pdq_cuda_one_frame_last_timing(double *out4, long long *kernel_clocks7);

/* ---- sharded small-needle query over peer memory (one process per GPU, no collective library) -------------------
 * The multi-GPU form of VulkanMatcher::scan (vulkan.rs:349-406): up to 32 needles on ONE rank, every rank scans its own
 * shard, the root learns which shards matched.  Every rank owns a small mailbox in its HBM that the others map through
 * CUDA IPC; needles and answers travel as plain stores over NVLink ordered by flags (csrc/api.cu), one stream
 * synchronisation on the root, none of NCCL's two collectives.
 *   create  : allocates this rank's mailbox; handle_out receives YUME_PDQ_CUDA_IPC_HANDLE_BYTES to hand to the peers
 *   connect : handles = world x YUME_PDQ_CUDA_IPC_HANDLE_BYTES, rank-ordered (exchanged by the caller, e.g. all-gather)
 *   scan    : COLLECTIVE -- every rank calls it once per query, in the same order, with the same q / threshold / root.
 *             The root passes the needles (host or device pointer), the others NULL.  On the root *matched_ranks receives
 *             a bit mask (bit r: rank r's shard holds a row within the threshold of some needle) and the call returns when
 *             the answer is in; on the other ranks the call only enqueues on `stream` (their answer travels to the root
 *             on the device) and *matched_ranks is 0.  YUME_PDQ_CUDA_E_TIMEOUT if a peer did not take part within 2 s.
 *             CHANGING THE ROOT: the root's call returns once every rank has answered, i.e. is done with its mailbox, and
 *             only the root learns that.  A different rank may therefore become the root only after the previous root's
 *             call has returned (any host-side ordering between the two processes, e.g. a barrier;
 *             yume_pdq_b200.sharded.ShardedMatcher.scan_from inserts one when the root changes) -- otherwise its needles
 *             could overwrite a mailbox that a slower rank, the previous root included, is still scanning from
 *             (tests/test_exchange_layout.py models both cases). */
#define YUME_PDQ_CUDA_IPC_HANDLE_BYTES 64
typedef struct YumePDQ_CudaMailbox YumePDQ_CudaMailbox;
int yume_pdq_cuda_mailbox_create(int device, int rank, int world, YumePDQ_CudaMailbox **out, uint8_t *handle_out);
int yume_pdq_cuda_mailbox_connect(YumePDQ_CudaMailbox *mb, const uint8_t *handles);
int yume_pdq_cuda_mailbox_destroy(YumePDQ_CudaMailbox *mb);
int yume_pdq_cuda_mailbox_scan(YumePDQ_CudaMailbox *mb, YumePDQ_CudaDb *db, const uint8_t *needles, size_t q,
                               uint32_t threshold, int root, uint32_t *matched_ranks, void *stream);
/* How a rank runs its part of a query.  AUTO: one launch that carries the whole exchange (needles out, wait, shard scan,
 * answer back, collection on the root) wherever the shard scan is the rows-in-M tensor kernel -- the kernel ALGO_AUTO
 * picks for shards of >= 1M rows at thresholds <= 40 -- and separate post / wait / scan / publish / collect launches
 * around ALGO_AUTO's kernel elsewhere.  SEPARATE / FUSED force one form (FUSED: any shard size, threshold <= 128).  The
 * two forms speak the same flag protocol: ranks may differ. */
#define YUME_PDQ_CUDA_MAILBOX_AUTO 0
#define YUME_PDQ_CUDA_MAILBOX_SEPARATE 1
#define YUME_PDQ_CUDA_MAILBOX_FUSED 2
int yume_pdq_cuda_mailbox_set_mode(YumePDQ_CudaMailbox *mb, int mode);
/* diagnostic: host-side microsecond stamps of the last query on the root (after the post, wait, scan, publish + collect
 * launches; answer in) */
int yume_pdq_cuda_mailbox_last_timing(const YumePDQ_CudaMailbox *mb, double *out5);

/* ---- needle all-gather over peer memory (one process per GPU, no collective library on the data path) ----
 * Every rank pushes its slice (slice_bytes, the same on every rank) into a buffer in every peer's HBM (CUDA IPC) followed by
 * a sequence number in the peer's flag word; `wait` makes a stream wait (one-warp kernel) until every peer's slice of the
 * latest `post` is in and returns this rank's buffer: the world - 1 slices of the OTHER ranks in rank order -- the packed
 * batch yume_pdq_cuda_matcher_find_device_split takes with split_at = rank * needles_per_slice and gap =
 * needles_per_slice.  How the slice travels: EXCHANGE_KERNEL -- one kernel stores it to all peers over NVLink and its last
 * CTA releases the flags (~15-20 us for 1 MiB to 7 peers; issue it BEFORE a persistent matcher kernel, not beside it);
 * EXCHANGE_COPY_ENGINE -- world - 1 peer copies + world - 1 four-byte flag copies in stream order (no launch of ours, but
 * ~15 us per copy, and issued beside a persistent kernel the copies delay it as an NCCL all-gather does).
 *   create / connect : as for the mailboxes (handle_out: YUME_PDQ_CUDA_IPC_HANDLE_BYTES; handles rank-ordered)
 *   post   : COLLECTIVE, enqueues on `stream`; slice_dev must stay unchanged until the stream has passed this point
 *   wait   : enqueues on `stream`; *others_dev is valid until the post after next (two buffers alternate).  The caller
 *            must keep ranks within one exchange of each other (a per-step collective such as the match-list gather does)
 *   failed : (synchronises) nonzero if a wait gave up after 2 s */
typedef struct YumePDQ_CudaExchange YumePDQ_CudaExchange;
int yume_pdq_cuda_exchange_create(int device, int rank, int world, size_t slice_bytes, YumePDQ_CudaExchange **out,
                                  uint8_t *handle_out);
int yume_pdq_cuda_exchange_connect(YumePDQ_CudaExchange *ex, const uint8_t *handles);
#define YUME_PDQ_CUDA_EXCHANGE_KERNEL 0
#define YUME_PDQ_CUDA_EXCHANGE_COPY_ENGINE 1
int yume_pdq_cuda_exchange_post(YumePDQ_CudaExchange *ex, const void *slice_dev, int how, void *stream);
/* a shorter slice this time (bytes <= slice_bytes, a multiple of 16, the same on every rank): the peers' buffers then hold
 * world - 1 slices of `bytes`, packed -- the last, partial chunk of a chunked pipeline */
int yume_pdq_cuda_exchange_post_bytes(YumePDQ_CudaExchange *ex, const void *slice_dev, size_t bytes, int how, void *stream);
int yume_pdq_cuda_exchange_wait(YumePDQ_CudaExchange *ex, void *stream, const uint8_t **others_dev);
int yume_pdq_cuda_exchange_failed(YumePDQ_CudaExchange *ex, int *failed);
int yume_pdq_cuda_exchange_destroy(YumePDQ_CudaExchange *ex);

/* Kernel selection for the matcher (0 = automatic) */
#define YUME_PDQ_CUDA_ALGO_AUTO 0
#define YUME_PDQ_CUDA_ALGO_EXACT 1   /* unfiltered: 8 POPC per (row, needle) pair */
#define YUME_PDQ_CUDA_ALGO_TILED 2   /* rows in registers x needles in smem, 96-bit carry-save prefilter */
/* 3 was a TMA-ring variant of ALGO_TILED (slower; removed) */
#define YUME_PDQ_CUDA_ALGO_TENSOR 4  /* tcgen05.mma kind::mxf4 on e2m1 operands: needle batches from 14 needles on. thresholds <= 36: 128-bit prefix filter on the tensor cores, two rows per accumulator (ALGO_TENSOR_PAIRED) + exact POPC check of the candidates; <= 127: full 256-bit dot products; > 127: kind::i8 */
#define YUME_PDQ_CUDA_ALGO_TENSOR_I8 5  /* tcgen05.mma kind::i8 on +-1 int8 expanded bits, any threshold */
#define YUME_PDQ_CUDA_ALGO_TENSOR_FULL 6    /* kind::mxf4, always the full 256-bit dot product (threshold <= 127) */
#define YUME_PDQ_CUDA_ALGO_TENSOR_PREFIX 7  /* kind::mxf4, always the 128-bit prefix filter (threshold <= 63) */
#define YUME_PDQ_CUDA_ALGO_TENSOR_PAIRED 8  /* kind::mxf4, 128-bit prefix filter with two rows per f32 accumulator (s16x2 fields; threshold <= 60) */
#define YUME_PDQ_CUDA_ALGO_TENSOR_ROWS 9    /* kind::mxf4 with the database rows in the M dimension (operand A in tensor memory): 1 .. 64 needles, 122-bit prefix filter + exact check */

/* PDQMatcher::scan (matching/mod.rs:72-79): *matched = 1 iff any needle is within
 * `threshold` (inclusive) of any row.  needles q x 32 bytes, host or device. */
int yume_pdq_cuda_matcher_scan(YumePDQ_CudaDb *db, const uint8_t *needles, size_t q,
                               uint32_t threshold, int *matched, void *stream);
/* PDQMatcher::find (matching/mod.rs:84-94) with the callback unrolled into a list:
 * every (row, needle) with distance <= threshold, sorted by (row, needle) -- the
 * order in which CpuMatcher::find (matching/mod.rs:336-362) invokes its callback.
 * At most `cap` entries are written to `out` (host memory); *n_found receives the
 * total; returns YUME_PDQ_CUDA_E_CAPACITY if it exceeds cap (the first cap entries
 * in sorted order are still returned). */
int yume_pdq_cuda_matcher_find(YumePDQ_CudaDb *db, const uint8_t *needles, size_t q,
                               uint32_t threshold, YumePDQ_Match *out, size_t cap,
                               size_t *n_found, void *stream);
/* VulkanMatcher::find semantics (vulkan.rs:408-472): the minimum of
 * needle + local_row * q over all matches, UINT64_MAX if none. */
int yume_pdq_cuda_matcher_find_min(YumePDQ_CudaDb *db, const uint8_t *needles, size_t q,
                                   uint32_t threshold, uint64_t *min_key, void *stream);
/* Fully device-side, asynchronous on `stream`: needles, out and count are DEVICE
 * pointers; *count (uint32) must be zeroed by the caller and receives the total
 * number of matches; the first `cap` (unsorted) are written to out.  `algo`
 * selects the kernel.  This is what bench.py times. */
int yume_pdq_cuda_matcher_find_device(YumePDQ_CudaDb *db, const uint8_t *needles_dev, size_t q,
                                      uint32_t threshold, YumePDQ_Match *out_dev, uint32_t cap,
                                      uint32_t *count_dev, int algo, void *stream);

/* Same, with `needle_base` added to every reported needle index: callers that match a long needle
 * stream in chunks (yume_pdq_b200/sharded.py pipelines chunked all-gathers against the local shard)
 * append to one list without a renumbering pass on the device. */
int yume_pdq_cuda_matcher_find_device_base(YumePDQ_CudaDb *db, const uint8_t *needles_dev, size_t q,
                                           uint32_t threshold, YumePDQ_Match *out_dev, uint32_t cap,
                                           uint32_t *count_dev, int algo, uint32_t needle_base, void *stream);
/* Same for a batch that is the concatenation of TWO ranges of a longer list with a hole between them -- every rank's
 * slice of an all-gathered needle batch but the caller's own, which it matched while the gather was in flight: needle j
 * is reported as needle_base + j for j < split_at and as needle_base + gap + j from split_at on. */
int yume_pdq_cuda_matcher_find_device_split(YumePDQ_CudaDb *db, const uint8_t *needles_dev, size_t q,
                                            uint32_t threshold, YumePDQ_Match *out_dev, uint32_t cap,
                                            uint32_t *count_dev, int algo, uint32_t needle_base, uint32_t split_at,
                                            uint32_t gap, void *stream);

/* Hash + match in one pipelined call: the flow of examples/end_to_end.rs:224-256 (hash_get_threshold ->
 * visit_dihedrals -> matcher.find for every frame) for a whole batch.  frames n x 512 x 512 f32, host or
 * device; host frames are copied in 128-frame chunks and chunk i's hash + match kernels overlap the copy
 * of chunk i+1.  Needle index of a match = 8 * frame + dihedral (YUME_PDQ_CUDA_DIHEDRAL_ORDER).
 * hashes (n x 8 x 32, nullable) and quality (n x 8, nullable) may be host or device; out (host, cap
 * entries) receives the matches sorted by (row, needle) exactly like yume_pdq_cuda_matcher_find;
 * *n_found the total; YUME_PDQ_CUDA_E_CAPACITY if it exceeds cap.  This is what bench.py's e2e leg times. */
int yume_pdq_cuda_screen_batch(YumePDQ_CudaDb *db, const float *frames, size_t n, uint32_t threshold,
                               uint8_t *hashes, float *quality, YumePDQ_Match *out, size_t cap,
                               size_t *n_found, void *stream);

/* yume_pdq_cuda_screen_batch for frames in any pixel format: with LUMA8 frames a step moves
 * 256 KiB per frame over PCIe instead of 1 MiB. */
int yume_pdq_cuda_screen_batch_pixels(YumePDQ_CudaDb *db, const void *frames, int pixel_format, size_t n,
                                      uint32_t threshold, uint8_t *hashes, float *quality,
                                      YumePDQ_Match *out, size_t cap, size_t *n_found, void *stream);

/* ---- measurement helper ----
 * Raw issue rate of the instruction the tensor matcher is built on (tcgen05.mma kind::mxf4.block_scale,
 * M 128, N 240, K 64, +-1.0 e2m1 operands, f32 accumulate) with no loads and no epilogue: one CTA per
 * SM issues `iters` x 8 MMAs back to back.  *tera_ops = 2 * M * N * K * MMAs / time / 1e12 (the roofline
 * denominator bench.py reports for that kernel); *sm_clk_per_mma (nullable) from clock64(). */
int yume_pdq_cuda_measure_mxf4_peak(int device, int iters, double *tera_ops, double *sm_clk_per_mma);

/* ---- synthetic inputs (device-side generators shared with the oracle) ----
 * frames_dev n x 512 x 512 f32 DEVICE pointer. kind 0: uniform 8-bit noise;
 * kind 1: smooth integer triangle waves + noise (see oracle/pdq_oracle.c). */
int yume_pdq_cuda_synth_frames(int device, float *frames_dev, size_t n, uint64_t first_frame,
                               uint64_t seed, int kind, void *stream);
/* 8-bit frames (LUMA8 / RGB8 / RGBA8), n x 512 x 512 x channels DEVICE bytes */
int yume_pdq_cuda_synth_pixels(int device, void *frames_dev, int pixel_format, size_t n,
                               uint64_t first_frame, uint64_t seed, void *stream);
int yume_pdq_cuda_synth_rows(int device, uint8_t *rows_dev, size_t n_rows, uint64_t first_row,
                             uint64_t seed, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* YUME_PDQ_CUDA_H */
