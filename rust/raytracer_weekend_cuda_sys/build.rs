//! Builds (or finds) librtw_cuda.so — the CUDA backend behind include/rtw_cuda.h — for sm_100a.
//!
//!   RTW_CUDA_LIB_DIR=<dir>   link the prebuilt library in <dir> (what `make -C raytracer-weekend_b200/csrc` leaves
//!                            in raytracer-weekend_b200/lib/)
//!   RTW_CUDA_SRC_DIR=<dir>   otherwise: compile <dir>/*.cu with nvcc (default: ../../raytracer-weekend_b200/csrc)
//!
//! -fmad=false is REQUIRED: the reference (rustc) never contracts a*b+c and the closest-hit parity is bit-exact.
use std::{env, path::PathBuf, process::Command};

fn main() {
    println!("cargo:rerun-if-env-changed=RTW_CUDA_LIB_DIR");
    println!("cargo:rerun-if-env-changed=RTW_CUDA_SRC_DIR");
    if let Ok(dir) = env::var("RTW_CUDA_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=rtw_cuda");
        return;
    }
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let src = env::var("RTW_CUDA_SRC_DIR")
        .map(PathBuf::from)
        .unwrap_or_else(|_| manifest.join("../../raytracer-weekend_b200/csrc"));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let units = ["rtw_api", "rtw_bvh", "rtw_trace", "rtw_render", "rtw_multi", "rtw_mem"];
    let mut objects = Vec::new();
    for u in units {
        let cu = src.join(format!("{u}.cu"));
        println!("cargo:rerun-if-changed={}", cu.display());
        let obj = out.join(format!("{u}.o"));
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false"])
            .args(["-Xcompiler", "-fPIC,-ffp-contract=off", "-c"])
            .arg(&cu)
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("nvcc not found: set NVCC or RTW_CUDA_LIB_DIR");
        assert!(status.success(), "nvcc failed on {}", cu.display());
        objects.push(obj);
    }
    for h in ["rtw_device.cuh", "rtw_scene.cuh", "rtw_traverse.cuh", "rtw_raysort.cuh"] {
        println!("cargo:rerun-if-changed={}", src.join(h).display());
    }
    let lib = out.join("librtw_cuda.so");
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o"])
        .arg(&lib)
        .args(&objects)
        .status()
        .expect("nvcc link failed to start");
    assert!(status.success(), "nvcc -shared failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=rtw_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
}
