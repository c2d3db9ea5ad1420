// build.rs -- compiles the CUDA library with nvcc for sm_100a and links it.
// (Authored, not compiled in this image: there is no Rust toolchain here.)
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let lib = out.join("libpairing_b200.a");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    // four translation units: the lane-pair pairing engine and the warp-cooperative engine are compiled on their own
    // (pairing_b200/csrc/abi_common.cuh); mgpu.cu is the host-side multi-device layer
    let mut objs = Vec::new();
    for unit in &["kernels", "kernels_pair", "kernels_mm", "kernels_wide", "mgpu"] {
        let src = root.join(format!("pairing_b200/csrc/{}.cu", unit));
        let obj = out.join(format!("{}.o", unit));
        let status = Command::new(&nvcc)
            .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                    "-Xcompiler", "-fPIC", "-c", "-o"])
            .arg(&obj)
            .arg(&src)
            .status()
            .expect("nvcc not found: set NVCC");
        assert!(status.success(), "nvcc failed");
        objs.push(obj);
    }
    let status = Command::new("ar").args(&["crs"]).arg(&lib).args(&objs).status().expect("ar");
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=pairing_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=pthread");
    for f in &["kernels.cu", "kernels_pair.cu", "kernels_mm.cu", "kernels_wide.cu", "mgpu.cu", "abi_common.cuh", "pair_io.cuh", "fr.cuh", "fp.cuh", "fp_inv_gcd.cuh", "fp_sqr_gen.cuh", "tower.cuh", "curve.cuh",
               "pair_tower.cuh", "wide.cuh", "wide_prog_gen.cuh", "codec.cuh", "constants.cuh"] {
        println!("cargo:rerun-if-changed={}", root.join("pairing_b200/csrc").join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/pairing_b200.h").display());
}
