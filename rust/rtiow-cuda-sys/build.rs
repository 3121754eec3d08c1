// Builds librtiow_cuda from the CUDA sources of this repository with nvcc, for sm_100a only
// (no multi-arch fatbin, no CPU fallback).  Set RTIOW_CUDA_LIB_DIR to link a prebuilt
// rtiow_b200/lib/librtiow_cuda.so instead of compiling.
//
// NOTE: written against cc 1.x; NOT compiled in the build container of this repository (no rustc/cargo there).
use std::{env, path::PathBuf};

fn main() {
    if let Ok(dir) = env::var("RTIOW_CUDA_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=rtiow_cuda");
        return;
    }
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("rtiow_b200/csrc");
    cc::Build::new()
        .cuda(true)
        .cudart("static")
        .flag("-gencode").flag("arch=compute_100a,code=sm_100a")
        .flag("-O3").flag("-lineinfo").flag("-std=c++17").flag("--expt-relaxed-constexpr")
        .include(root.join("include"))
        .file(csrc.join("capi.cu"))
        .file(csrc.join("scene_gen.cpp"))
        .compile("rtiow_cuda");
    for f in ["capi.cu", "scene_gen.cpp", "rt_device.cuh", "rt_scene.cuh", "rt_render.cuh", "rt_unit.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/rtiow_cuda.h").display());
}
