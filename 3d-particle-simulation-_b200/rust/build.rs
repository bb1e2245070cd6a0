// build.rs — compiles the CUDA sources of the B200 step engine with nvcc for sm_100a and links the
// resulting static library into the crate (north_star: "A build.rs compiles the .cu sources with
// nvcc -arch=sm_100a").  The reference crate has no build script (Cargo.toml:1-13).
//
// Expects the engine sources next to the crate:  <crate>/p3d/{include/p3d.h, csrc/*.cu, csrc/*.cuh, csrc/*.cpp}
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("p3d");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let cuda_lib = env::var("CUDA_LIB_DIR").unwrap_or_else(|_| "/usr/local/cuda/lib64".into());
    let srcs = ["csrc/p3d_engine.cu", "csrc/p3d_scene.cpp"];
    let mut objs = Vec::new();
    for s in srcs {
        let obj = out.join(format!("{}.o", s.replace('/', "_")));
        let ok = Command::new(&nvcc)
            .args(["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a"])
            .args(["-Xcompiler", "-fPIC", "-c"])
            .arg(format!("-I{}", root.join("include").display()))
            .arg(format!("-I{}", root.join("csrc").display()))
            .arg("-o").arg(&obj).arg(root.join(s))
            .status().expect("nvcc not found").success();
        assert!(ok, "nvcc failed on {s}");
        objs.push(obj);
        println!("cargo:rerun-if-changed={}", root.join(s).display());
    }
    let lib = out.join("libp3d.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-search=native={cuda_lib}");
    println!("cargo:rustc-link-lib=static=p3d");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
