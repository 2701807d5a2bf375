//! CUDA matcher next to matching/vulkan.rs.  NOT COMPILED in the build image.
use generic_array::{ArrayLength, GenericArray, sequence::Flatten, typenum::{B1, U32}};
use super::PDQMatcher;
use crate::{cuda_sys as sys, kernel::cuda::CudaAccelerator};

/// A vector database already uploaded to the accelerator (mirrors VulkanVectorDatabase::new, vulkan.rs:50-64).
pub struct CudaVectorDatabase { db: *mut sys::YumePDQ_CudaDb }
unsafe impl Send for CudaVectorDatabase {}

impl CudaVectorDatabase {
    pub fn new(device: i32, data: &[GenericArray<u8, U32>], global_row_offset: u64) -> Self {
        let mut db = core::ptr::null_mut();
        let rc = unsafe { sys::yume_pdq_cuda_db_create(device, data.as_ptr().cast(), data.len(), global_row_offset, &mut db) };
        assert!(rc == 0, "yume_pdq_cuda_db_create: {rc}");
        Self { db }
    }
    pub fn len(&self) -> usize { unsafe { sys::yume_pdq_cuda_db_len(self.db) } }
}
impl Drop for CudaVectorDatabase {
    fn drop(&mut self) { unsafe { sys::yume_pdq_cuda_db_destroy(self.db); } }
}

pub struct CudaMatcher<'b, NumNeedles: ArrayLength> {
    haystack: &'b CudaVectorDatabase,
    threshold: u32,
    _marker: core::marker::PhantomData<NumNeedles>,
}

impl<'b, NumNeedles: ArrayLength> CudaMatcher<'b, NumNeedles> {
    pub fn new(haystack: &'b CudaVectorDatabase, threshold: u32) -> Self {
        Self { haystack, threshold, _marker: core::marker::PhantomData }
    }
}

impl<NumNeedles: ArrayLength> PDQMatcher for CudaMatcher<'_, NumNeedles> {
    type BatchSize = NumNeedles;
    type RequiredHardwareFeature = CudaAccelerator;
    type InputDimension = U32;
    type Aligner<T> = T;
    type NoFalseNegative = B1;
    type NoFalsePositive = B1;
    type FindsAllMatches = B1;     // unlike VulkanMatcher (B0, vulkan.rs:343)
    type Ident = &'static str;
    fn ident() -> Self::Ident { "CudaMatcher (sm_100a)" }

    fn scan(&mut self, query: &GenericArray<GenericArray<u8, U32>, NumNeedles>) -> bool {
        let mut m = 0;
        let rc = unsafe { sys::yume_pdq_cuda_matcher_scan(self.haystack.db, query.as_ptr().cast(), NumNeedles::USIZE,
            self.threshold, &mut m, core::ptr::null_mut()) };
        assert!(rc == 0, "yume_pdq_cuda_matcher_scan: {rc}");
        m != 0
    }

    fn find<R, F: FnMut(usize, usize) -> Option<R>>(&mut self,
            query: &GenericArray<GenericArray<u8, U32>, NumNeedles>, mut f: F) -> Option<R> {
        let mut cap = 1024usize;
        loop {
            let mut out = vec![sys::YumePDQ_Match::default(); cap];
            let mut n = 0usize;
            let rc = unsafe { sys::yume_pdq_cuda_matcher_find(self.haystack.db, query.as_ptr().cast(), NumNeedles::USIZE,
                self.threshold, out.as_mut_ptr(), cap, &mut n, core::ptr::null_mut()) };
            if rc == sys::E_CAPACITY { cap = n; continue; }
            assert!(rc == 0, "yume_pdq_cuda_matcher_find: {rc}");
            // (row, needle) lexicographic order == CpuMatcher::find's callback order (matching/mod.rs:336-362)
            for m in &out[..n] {
                if let Some(r) = f(m.row as usize, m.needle as usize) { return Some(r); }
            }
            return None;
        }
    }
}
