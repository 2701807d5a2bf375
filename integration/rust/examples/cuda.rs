//! `cargo run --release --features cuda --example cuda` -- the CUDA twin of examples/vulkan.rs:148-216.  NOT COMPILED in
//! the build image (no rustc); integration/cpp/example.cpp runs the same flow through the same C ABI and IS tested.
//!
//! 10M random rows resident on GPU 0, one row replaced by a near-duplicate of the test image's hash; hash the image
//! with the CUDA kernel, collect the 8 dihedral hashes, find the planted row, then time a scan that finds nothing.
use generic_array::{GenericArray, sequence::Flatten, typenum::{U8, U32}};
use rand::RngCore;
use yume_pdq::{
    kernel::{Kernel, cuda::CudaKernel},
    matching::{PDQMatcher, cuda::{CudaMatcher, CudaVectorDatabase}},
};

fn main() {
    const N: usize = 10_000_000;
    let mut rng = rand::rng();
    let mut rows = vec![GenericArray::<u8, U32>::default(); N];
    rows.iter_mut().for_each(|r| rng.fill_bytes(r.as_mut_slice()));

    // hash the test image (512 x 512 luma as f32) and its 7 other dihedrals
    let image = image::open("test-data/aaa-orig.jpg").unwrap()
        .resize_exact(512, 512, image::imageops::FilterType::Triangle).to_luma8();
    let mut pixels = GenericArray::<GenericArray<f32, _>, _>::default();
    for (y, row) in pixels.iter_mut().enumerate() {
        for (x, p) in row.iter_mut().enumerate() { *p = image.get_pixel(x as u32, y as u32).0[0] as f32; }
    }
    let mut kernel = CudaKernel::new(0);
    let (mut threshold, mut hash, mut buf1, mut pdqf) = Default::default();
    yume_pdq::hash_get_threshold(&mut kernel, &pixels, &mut threshold, &mut hash, &mut buf1,
                                 &mut GenericArray::default(), &mut pdqf);
    let mut dihedrals = GenericArray::<GenericArray<u8, U32>, U8>::default();
    dihedrals[0] = hash.flatten().clone();
    let mut i = 1;
    yume_pdq::visit_dihedrals(&mut kernel, &mut pdqf, &mut hash, threshold, |_d, _i, (_, _, h)| {
        dihedrals[i] = h.flatten().clone();
        i += 1;
        Ok::<(), ()>(())
    }).unwrap();

    // plant a copy of dihedral 3, 12 bits away
    let insertion_index = (rng.next_u64() as usize) % N;
    rows[insertion_index] = dihedrals[3].clone();
    for b in 0..12 { rows[insertion_index][b] ^= 1 << (b % 8); }

    let haystack = CudaVectorDatabase::new(0, &rows, 0);         // one upload; the rows stay in HBM
    let mut matcher = CudaMatcher::<U8>::new(&haystack, 31);
    let found = matcher.find(&dihedrals, |row, needle| Some((row, needle)));
    assert_eq!(found, Some((insertion_index, 3)));

    dihedrals.iter_mut().for_each(|h| rng.fill_bytes(h.as_mut_slice()));
    let start = std::time::Instant::now();
    assert!(!matcher.scan(&dihedrals), "found a match in random data");
    println!("10M-row scan, 8 needles: {:?} (host call incl. 256 B upload + 16 B download)", start.elapsed());
}
