//! Raw bindings of include/yume_pdq_cuda.h (group B).  NOT COMPILED in the build image.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
pub struct YumePDQ_CudaDb { _private: [u8; 0] }

#[repr(C)]
pub struct YumePDQ_CudaMailbox { _private: [u8; 0] }
#[repr(C)]
pub struct YumePDQ_CudaExchange { _private: [u8; 0] }
pub const IPC_HANDLE_BYTES: usize = 64;

#[repr(C)]
#[derive(Debug, Clone, Copy, PartialEq, Eq, Default)]
pub struct YumePDQ_Match { pub row: u64, pub needle: u32, pub dist: u32 }

pub const E_CAPACITY: c_int = -3;
pub const PIXEL_F32: c_int = 0;
pub const PIXEL_LUMA8: c_int = 1;
pub const PIXEL_RGB8: c_int = 2;
pub const PIXEL_RGBA8: c_int = 3;
pub const PIXEL_FMA: c_int = 0x100;

unsafe extern "C" {
    pub fn yume_pdq_cuda_error_string(code: c_int) -> *const c_char;
    pub fn yume_pdq_cuda_device_count() -> c_int;
    pub fn yume_pdq_cuda_kernel_ident() -> *const c_char;
    pub fn yume_pdq_cuda_hash_batch(device: c_int, frames: *const f32, n: usize, hashes: *mut u8,
        quality: *mut f32, thresholds: *mut f32, pdqf: *mut f32, stream: *mut c_void) -> c_int;
    pub fn yume_pdq_cuda_hash_batch_pixels(device: c_int, frames: *const c_void, pixel_format: c_int, n: usize,
        hashes: *mut u8, quality: *mut f32, thresholds: *mut f32, pdqf: *mut f32, stream: *mut c_void) -> c_int;
    pub fn yume_pdq_cuda_cvt_to_luma8f(device: c_int, pixels: *const c_void, pixel_format: c_int, n_pixels: usize,
        coeff: *const f32, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn yume_pdq_cuda_host_alloc(device: c_int, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn yume_pdq_cuda_host_free(p: *mut c_void) -> c_int;
    pub fn yume_pdq_cuda_jarosz_compress(device: c_int, frames: *const f32, n: usize, buf1: *mut f32,
        stream: *mut c_void) -> c_int;
    pub fn yume_pdq_cuda_dct2d(device: c_int, buf1: *const f32, n: usize, pdqf: *mut f32,
        stream: *mut c_void) -> c_int;
    pub fn yume_pdq_cuda_sum_of_gradients(device: c_int, pdqf: *const f32, n: usize, gradient: *mut f32,
        quality: *mut f32, stream: *mut c_void) -> c_int;
    pub fn yume_pdq_cuda_quantize(device: c_int, pdqf: *const f32, n: usize, thresholds: *mut f32,
        hashes: *mut u8, stream: *mut c_void) -> c_int;
    pub fn yume_pdq_cuda_pdqf_transform(device: c_int, pdqf: *mut f32, n: usize, op: c_int,
        stream: *mut c_void) -> c_int;
    pub fn yume_pdq_cuda_db_create(device: c_int, rows: *const u8, n_rows: usize, global_row_offset: u64,
        out: *mut *mut YumePDQ_CudaDb) -> c_int;
    pub fn yume_pdq_cuda_db_destroy(db: *mut YumePDQ_CudaDb) -> c_int;
    pub fn yume_pdq_cuda_db_len(db: *const YumePDQ_CudaDb) -> usize;
    pub fn yume_pdq_cuda_matcher_scan(db: *mut YumePDQ_CudaDb, needles: *const u8, q: usize, threshold: u32,
        matched: *mut c_int, stream: *mut c_void) -> c_int;
    pub fn yume_pdq_cuda_matcher_find(db: *mut YumePDQ_CudaDb, needles: *const u8, q: usize, threshold: u32,
        out: *mut YumePDQ_Match, cap: usize, n_found: *mut usize, stream: *mut c_void) -> c_int;
    /// hash + 8 dihedrals + match for a whole batch (examples/end_to_end.rs:224-256 in one pipelined call)
    pub fn yume_pdq_cuda_screen_batch_pixels(db: *mut YumePDQ_CudaDb, frames: *const c_void, pixel_format: c_int,
        n: usize, threshold: u32, hashes: *mut u8, quality: *mut f32, out: *mut YumePDQ_Match, cap: usize,
        n_found: *mut usize, stream: *mut c_void) -> c_int;
    // ---- one process per GPU (sharded database): device-side scans, peer-memory mailboxes, needle exchange ----
    pub fn yume_pdq_cuda_matcher_find_device_split(db: *mut YumePDQ_CudaDb, needles_dev: *const u8, q: usize,
        threshold: u32, out_dev: *mut YumePDQ_Match, cap: u32, count_dev: *mut u32, algo: c_int, needle_base: u32,
        split_at: u32, gap: u32, stream: *mut c_void) -> c_int;
    pub fn yume_pdq_cuda_mailbox_create(device: c_int, rank: c_int, world: c_int, out: *mut *mut YumePDQ_CudaMailbox,
        handle_out: *mut u8) -> c_int;
    pub fn yume_pdq_cuda_mailbox_connect(mb: *mut YumePDQ_CudaMailbox, handles: *const u8) -> c_int;
    pub fn yume_pdq_cuda_mailbox_set_mode(mb: *mut YumePDQ_CudaMailbox, mode: c_int) -> c_int;
    pub fn yume_pdq_cuda_mailbox_scan(mb: *mut YumePDQ_CudaMailbox, db: *mut YumePDQ_CudaDb, needles: *const u8, q: usize,
        threshold: u32, root: c_int, matched_ranks: *mut u32, stream: *mut c_void) -> c_int;
    pub fn yume_pdq_cuda_mailbox_destroy(mb: *mut YumePDQ_CudaMailbox) -> c_int;
    pub fn yume_pdq_cuda_exchange_create(device: c_int, rank: c_int, world: c_int, slice_bytes: usize,
        out: *mut *mut YumePDQ_CudaExchange, handle_out: *mut u8) -> c_int;
    pub fn yume_pdq_cuda_exchange_connect(ex: *mut YumePDQ_CudaExchange, handles: *const u8) -> c_int;
    pub fn yume_pdq_cuda_exchange_post(ex: *mut YumePDQ_CudaExchange, slice_dev: *const c_void, how: c_int,
        stream: *mut c_void) -> c_int;
    pub fn yume_pdq_cuda_exchange_wait(ex: *mut YumePDQ_CudaExchange, stream: *mut c_void,
        others_dev: *mut *const u8) -> c_int;
    pub fn yume_pdq_cuda_exchange_failed(ex: *mut YumePDQ_CudaExchange, failed: *mut c_int) -> c_int;
    pub fn yume_pdq_cuda_exchange_destroy(ex: *mut YumePDQ_CudaExchange) -> c_int;
}
