// Link step for the crate that hosts ola_gpu.rs (plonky2/plonky2/build.rs in the reference; the field crate's
// build.rs:4-27, which links the source-less libcuda_lib.a, is dropped together with the `cuda` feature).
fn main() {
    let dir = std::env::var("OLA_GPU_LIB_DIR").expect("set OLA_GPU_LIB_DIR to the directory holding libola_gpu.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=ola_gpu");
    println!("cargo:rerun-if-env-changed=OLA_GPU_LIB_DIR");
}
