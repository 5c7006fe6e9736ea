// build.rs -- compiles ../csrc/*.cu for sm_100a with nvcc into a static library and links it plus cudart.
// (Mirrors rust-la_b200/Makefile, which is what the test harness uses because cargo is unavailable in the image.)
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".to_string());
    let nvcc = format!("{}/bin/nvcc", cuda);
    let srcs = ["la_runtime", "gemm_f64", "gemm_f32", "gemm_simt", "lu", "lu_solve", "cholesky", "qr", "elementwise", "mg",
                "capi"];
    let mut objs = Vec::new();
    for s in srcs.iter() {
        let src = format!("../csrc/{}.cu", s);
        let obj = out.join(format!("{}.o", s));
        println!("cargo:rerun-if-changed={}", src);
        let st = Command::new(&nvcc)
            .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                    "-Xcompiler", "-fPIC", "-c", &src, "-o"])
            .arg(&obj)
            .status()
            .expect("nvcc not found: the la crate has no CPU fallback and needs the CUDA toolkit");
        assert!(st.success(), "nvcc failed on {}", src);
        objs.push(obj);
    }
    println!("cargo:rerun-if-changed=../csrc/la_common.cuh");
    println!("cargo:rerun-if-changed=../csrc/ll_exchange.cuh");
    println!("cargo:rerun-if-changed=../../include/la_cabi.h");
    let lib = out.join("libla_b200.a");
    let st = Command::new("ar").arg("crs").arg(&lib).args(&objs).status().expect("ar");
    assert!(st.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-search=native={}/lib64", cuda);
    println!("cargo:rustc-link-lib=static=la_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=pthread");
}
