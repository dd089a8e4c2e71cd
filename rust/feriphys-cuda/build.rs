// build.rs -- compiles the CUDA sources for sm_100a with nvcc and links the result.
// (feriphys's own build.rs only copies assets, build.rs:1-18; this one has no ancestor.)
use std::{env, path::PathBuf, process::Command};

const SOURCES: &[&str] = &[
    "fp_api.cu", "fp_allpairs.cu", "fp_small.cu", "fp_grid.cu", "fp_walk.cu", "fp_sort.cu",
    "fp_misc.cu", "fp_shard.cu",
];

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    if env::var("CARGO_FEATURE_PREBUILT").is_ok() {
        let dir = env::var("FERIPHYS_CUDA_LIB_DIR").expect("set FERIPHYS_CUDA_LIB_DIR");
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=feriphys_cuda");
        return;
    }
    // repository layout: <root>/rust/feriphys-cuda/build.rs, <root>/feriphys_b200/csrc/*.cu
    let csrc = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../feriphys_b200/csrc");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let lib = out.join("libferiphys_cuda.so");
    let mut cmd = Command::new(nvcc);
    cmd.args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "--cudart", "static", "-shared", "-Xcompiler", "-fPIC", "-o"])
        .arg(&lib);
    for s in SOURCES {
        let p = csrc.join(s);
        println!("cargo:rerun-if-changed={}", p.display());
        cmd.arg(p);
    }
    cmd.arg("-ldl");
    let status = cmd.status().expect("failed to run nvcc (CUDA 12.8+ with sm_100a support is required)");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=feriphys_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
}
