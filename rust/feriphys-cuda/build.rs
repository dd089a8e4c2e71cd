// build.rs -- compiles the CUDA sources for sm_100a with nvcc and links the result.
// (feriphys's own build.rs only copies assets, build.rs:1-18; this one has no ancestor.)
//
// The list of translation units is READ from feriphys_b200/csrc/Makefile (its SRCS line), so the
// crate and the in-tree build cannot drift apart (tests/test_rust_crate.py checks the parse).
use std::{env, fs, path::PathBuf, process::Command};

fn makefile_sources(makefile: &str) -> Vec<String> {
    let text = fs::read_to_string(makefile).expect("cannot read feriphys_b200/csrc/Makefile");
    let line = text
        .lines()
        .find(|l| l.trim_start().starts_with("SRCS"))
        .expect("no SRCS line in the Makefile");
    line.split_once('=')
        .expect("malformed SRCS line")
        .1
        .split_whitespace()
        .filter(|w| w.ends_with(".cu"))
        .map(str::to_owned)
        .collect()
}

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    if env::var("CARGO_FEATURE_PREBUILT").is_ok() {
        let dir = env::var("FERIPHYS_CUDA_LIB_DIR").expect("set FERIPHYS_CUDA_LIB_DIR");
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=feriphys_cuda");
        return;
    }
    // repository layout: <root>/rust/feriphys-cuda/build.rs, <root>/feriphys_b200/csrc/*.cu
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("feriphys_b200/csrc");
    let makefile = csrc.join("Makefile");
    println!("cargo:rerun-if-changed={}", makefile.display());
    let sources = makefile_sources(makefile.to_str().unwrap());
    assert!(!sources.is_empty(), "the Makefile lists no .cu sources");
    // every header the translation units include
    for entry in fs::read_dir(&csrc).unwrap().flatten() {
        let p = entry.path();
        if matches!(p.extension().and_then(|e| e.to_str()), Some("cuh") | Some("h")) {
            println!("cargo:rerun-if-changed={}", p.display());
        }
    }
    println!("cargo:rerun-if-changed={}", root.join("include/feriphys_cuda.h").display());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let lib = out.join("libferiphys_cuda.so");
    let mut cmd = Command::new(nvcc);
    cmd.args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "--cudart", "static", "-shared", "-Xcompiler", "-fPIC", "-o"])
        .arg(&lib);
    for s in &sources {
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
