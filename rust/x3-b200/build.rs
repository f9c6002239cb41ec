// build.rs -- compiles the hand-written sm_100a kernels and the C ABI with nvcc and links them.
// (north_star: "The host side stays in Rust and calls hand-written sm_100a CUDA through a thin extern "C" FFI
//  built by build.rs (nvcc, no Triton, no CPU fallback)".)  NOT exercised in this repository: no Rust toolchain.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../../x3-rust_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libx3b200.a");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut objs = Vec::new();
    for src in ["x3_api.cu", "x3_encode.cu", "x3_decode.cu", "x3_synth.cu"] {
        let obj = out.join(src.replace(".cu", ".o"));
        let ok = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                   "-Xcompiler", "-fPIC", "-c", "-o"])
            .arg(&obj)
            .arg(csrc.join(src))
            .status()
            .expect("nvcc not found")
            .success();
        assert!(ok, "nvcc failed on {src}");
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
        objs.push(obj);
    }
    let ok = Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success();
    assert!(ok);
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=x3b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=static=cudart_static");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rustc-link-lib=dylib=rt");
    println!("cargo:rustc-link-lib=dylib=pthread");
}
