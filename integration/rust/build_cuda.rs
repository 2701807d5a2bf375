// Appended to the reference's build.rs, called from main() when CARGO_FEATURE_CUDA is set.
// Compiles the CUDA sources for sm_100a with nvcc and links the resulting static library.
#[cfg(feature = "cuda")]
fn build_cuda() {
    use std::{env, path::PathBuf, process::Command};
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let src = PathBuf::from("cuda/csrc"); // yume_pdq_b200/csrc of this repository
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let lib = out.join("libyume_pdq_cuda.a");
    let mut objs = vec![];
    // every translation unit of yume_pdq_b200/csrc (kept in step with SOURCES in yume_pdq_b200/build.py;
    // integration/rust/check.sh fails when the two lists differ)
    for f in ["api.cu", "hasher.cu", "matcher.cu", "matcher_tc.cu", "matcher_tc4.cu", "matcher_tc4x2.cu", "matcher_tc1.cu",
              "synth.cu", "peak.cu"] {
        let obj = out.join(f).with_extension("o");
        let ok = Command::new(&nvcc)
            .args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                   "-Xcompiler", "-fPIC", "-I", "cuda/include", "-c"])
            .arg(src.join(f)).arg("-o").arg(&obj)
            .status().expect("nvcc not found").success();
        assert!(ok, "nvcc failed on {f}");
        println!("cargo:rerun-if-changed={}", src.join(f).display());
        objs.push(obj);
    }
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=yume_pdq_cuda");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
