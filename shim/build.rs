// build.rs -- compile the CUDA side with nvcc for sm_100a (the same command as the repository's Makefile and
// bls_bn254_b200/_native.py) and link it.  Replaces nothing in the reference (it has no build script).
use std::{env, path::PathBuf, process::Command};

fn main() {
    println!("cargo:rerun-if-env-changed=BLS_BN254_B200_LIB_DIR");
    println!("cargo:rerun-if-env-changed=NVCC");
    if let Ok(dir) = env::var("BLS_BN254_B200_LIB_DIR") {
        // prebuilt library
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=bls_bn254_b200");
        return;
    }
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let src = root.join("bls_bn254_b200/csrc/capi.cu");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let so = out.join("libbls_bn254_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--shared", "-Xcompiler", "-fPIC", "-o"])
        .arg(&so)
        .arg(&src)
        .status()
        .unwrap_or_else(|e| panic!("cannot run {nvcc}: {e} (CUDA 12.8+ with sm_100a support is required; there is no CPU fallback)"));
    assert!(status.success(), "nvcc failed on {}", src.display());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=bls_bn254_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
    println!("cargo:rerun-if-changed={}", root.join("bls_bn254_b200/csrc").display());
    println!("cargo:rerun-if-changed={}", root.join("include/bls_bn254_b200.h").display());
}
