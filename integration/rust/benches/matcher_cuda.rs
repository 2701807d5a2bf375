// Goes into benches/matcher.rs next to the `vulkan` block (benches/matcher.rs:85-140), inside the
// "8needles/scan10mil" group.  NOT COMPILED in the build image.
#[cfg(feature = "cuda")]
{
    use yume_pdq::matching::cuda::{CudaMatcher, CudaVectorDatabase};
    // the whole 10M-row haystack is ONE resident database (the Vulkan block needs two: 160 MB binding limit)
    let haystack = CudaVectorDatabase::new(0, unsafe {
        core::slice::from_raw_parts::<GenericArray<u8, U32>>(batches[0].as_ptr().cast(), NUM_BATCHES * 2 * 1024)
    }, 0);
    let mut cuda = CudaMatcher::<U8>::new(&haystack, 31);
    group.bench_function("cuda", |b| {
        b.iter(|| assert!(!cuda.scan(&needles_data)));
    });
}
