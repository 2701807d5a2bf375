//! `impl Kernel for CudaKernel` -- sits next to x86.rs / portable_simd.rs.  NOT COMPILED in the build image.
//!
//! The per-frame trait methods each make one (synchronous, host-pointer) call into the batched C ABI
//! with n = 1; throughput users call `CudaKernel::hash_batch` directly.
use generic_array::{GenericArray, typenum::{B1, U2, U16, U128, U512}};
use super::{Kernel, type_traits::{EvaluateHardwareFeature, kernel_sealing::KernelSealed}};
use crate::cuda_sys as sys;

/// Hardware-feature marker like `VulkanAccelerator` (matching/vulkan.rs:295-333).
#[derive(Debug, Clone, Copy)]
pub struct CudaAccelerator;
impl KernelSealed for CudaAccelerator {}
impl EvaluateHardwareFeature for CudaAccelerator {
    type Name = &'static str;
    type EnabledStatic = B1;
    type MustCheck = B1;
    fn name() -> Self::Name { "cuda_sm100" }
    fn met_runtime() -> bool { unsafe { sys::yume_pdq_cuda_device_count() > 0 } }
}

#[derive(Clone, Copy, Default)]
pub struct CudaKernel { pub device: i32 }

fn ok(rc: i32, what: &str) { assert!(rc == 0, "{what}: yume_pdq_cuda error {rc}"); }

impl CudaKernel {
    /// Fused batched hot path: n frames -> n x 8 hashes (dihedral order of visit_dihedrals, NORMAL first).
    pub fn hash_batch(&self, frames: &[f32], hashes: &mut [u8], quality: &mut [f32], thresholds: &mut [f32]) {
        let n = frames.len() / (512 * 512);
        assert!(hashes.len() == n * 256 && quality.len() == n * 8 && thresholds.len() == n * 8);
        ok(unsafe { sys::yume_pdq_cuda_hash_batch(self.device, frames.as_ptr(), n, hashes.as_mut_ptr(),
            quality.as_mut_ptr(), thresholds.as_mut_ptr(), core::ptr::null_mut(), core::ptr::null_mut()) }, "hash_batch");
    }
    /// The same for raw 8-bit frames (`sys::PIXEL_LUMA8 / RGB8 / RGBA8`): the input conversion of the CLI
    /// (main.rs:486-608) runs inside the hash kernel, a frame crosses PCIe as bytes.
    pub fn hash_batch_pixels(&self, frames: &[u8], pixel_format: i32, hashes: &mut [u8], quality: &mut [f32],
                             thresholds: &mut [f32]) {
        let channels = match pixel_format & 0xff { sys::PIXEL_LUMA8 => 1, sys::PIXEL_RGB8 => 3, _ => 4 };
        let n = frames.len() / (512 * 512 * channels);
        assert!(hashes.len() == n * 256 && quality.len() == n * 8 && thresholds.len() == n * 8);
        ok(unsafe { sys::yume_pdq_cuda_hash_batch_pixels(self.device, frames.as_ptr().cast(), pixel_format, n,
            hashes.as_mut_ptr(), quality.as_mut_ptr(), thresholds.as_mut_ptr(), core::ptr::null_mut(),
            core::ptr::null_mut()) }, "hash_batch_pixels");
    }
    fn cvt(&self, pixels: *const core::ffi::c_void, fmt: i32, n: usize, coeff: [f32; 3], out: *mut f32) {
        ok(unsafe { sys::yume_pdq_cuda_cvt_to_luma8f(self.device, pixels, fmt, n, coeff.as_ptr(), out,
            core::ptr::null_mut()) }, "cvt_to_luma8f");
    }
    fn transform(&self, p: &mut GenericArray<GenericArray<f32, U16>, U16>, op: i32) {
        ok(unsafe { sys::yume_pdq_cuda_pdqf_transform(self.device, p.as_mut_ptr().cast(), 1, op, core::ptr::null_mut()) }, "pdqf_transform");
    }
}

impl Kernel for CudaKernel {
    type Buffer1WidthX = U128;   // must equal U128 to compose with the x86 kernels in a KernelRouter (router.rs:433-446)
    type Buffer1LengthY = U128;
    type InternalFloat = f32;
    type InputDimension = U512;
    type OutputDimension = U16;
    type RequiredHardwareFeature = CudaAccelerator;
    type Ident = &'static str;

    fn ident(&self) -> &'static str { "cuda_sm100a_f32_bitexact_scalar_order" }
    fn adjust_quality(input: f32) -> f32 { (input / super::QUALITY_ADJUST_DIVISOR as f32).min(1.0) }

    // mod.rs:303-340: the coefficients arrive as f32 bit patterns in const generics
    fn cvt_rgb8_to_luma8f<const R_COEFF: u32, const G_COEFF: u32, const B_COEFF: u32>(
        &mut self, input: &GenericArray<GenericArray<u8, generic_array::typenum::U3>, U512>,
        output: &mut GenericArray<f32, U512>) {
        self.cvt(input.as_ptr().cast(), sys::PIXEL_RGB8, 512,
                 [f32::from_bits(R_COEFF), f32::from_bits(G_COEFF), f32::from_bits(B_COEFF)], output.as_mut_ptr());
    }
    fn cvt_rgba8_to_luma8f<const R_COEFF: u32, const G_COEFF: u32, const B_COEFF: u32>(
        &mut self, input: &GenericArray<GenericArray<u8, generic_array::typenum::U4>, U512>,
        output: &mut GenericArray<f32, U512>) {
        self.cvt(input.as_ptr().cast(), sys::PIXEL_RGBA8, 512,
                 [f32::from_bits(R_COEFF), f32::from_bits(G_COEFF), f32::from_bits(B_COEFF)], output.as_mut_ptr());
    }

    fn jarosz_compress(&mut self, buffer: &GenericArray<GenericArray<f32, U512>, U512>,
                       output: &mut GenericArray<GenericArray<f32, U128>, U128>) {
        ok(unsafe { sys::yume_pdq_cuda_jarosz_compress(self.device, buffer.as_ptr().cast(), 1,
            output.as_mut_ptr().cast(), core::ptr::null_mut()) }, "jarosz_compress");
    }
    fn dct2d(&mut self, buffer: &GenericArray<GenericArray<f32, U128>, U128>, _tmp: &mut GenericArray<f32, U128>,
             output: &mut GenericArray<GenericArray<f32, U16>, U16>) {
        ok(unsafe { sys::yume_pdq_cuda_dct2d(self.device, buffer.as_ptr().cast(), 1, output.as_mut_ptr().cast(),
            core::ptr::null_mut()) }, "dct2d");
    }
    fn sum_of_gradients(&mut self, input: &GenericArray<GenericArray<f32, U16>, U16>) -> f32 {
        let mut g = 0f32;
        ok(unsafe { sys::yume_pdq_cuda_sum_of_gradients(self.device, input.as_ptr().cast(), 1, &mut g,
            core::ptr::null_mut(), core::ptr::null_mut()) }, "sum_of_gradients");
        g
    }
    fn quantize(&mut self, input: &GenericArray<GenericArray<f32, U16>, U16>, threshold: &mut f32,
                output: &mut GenericArray<GenericArray<u8, U2>, U16>) {
        ok(unsafe { sys::yume_pdq_cuda_quantize(self.device, input.as_ptr().cast(), 1, threshold,
            output.as_mut_ptr().cast(), core::ptr::null_mut()) }, "quantize");
    }
    fn pdqf_t(&mut self, input: &mut GenericArray<GenericArray<f32, U16>, U16>) { self.transform(input, 0) }
    fn pdqf_negate_alt_cols<const NEGATE: bool>(&mut self, input: &mut GenericArray<GenericArray<f32, U16>, U16>) {
        self.transform(input, if NEGATE { 2 } else { 1 })
    }
    fn pdqf_negate_alt_rows<const NEGATE: bool>(&mut self, input: &mut GenericArray<GenericArray<f32, U16>, U16>) {
        self.transform(input, if NEGATE { 3 } else { 4 })   // AVX2 convention: NEGATE=true starts at row 0
    }
    fn pdqf_negate_off_diagonals(&mut self, input: &mut GenericArray<GenericArray<f32, U16>, U16>) {
        self.transform(input, 5)
    }
}
