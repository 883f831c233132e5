// UNVERIFIED SOURCE (no Rust toolchain in the build image).
// Links libpedoni_cuda.so, built by `python -m pedoni_b200.build` (nvcc, sm_100a) in the pedoni-b200 repo.
fn main() {
    let dir = std::env::var("PEDONI_CUDA_LIB_DIR").expect("set PEDONI_CUDA_LIB_DIR to the directory holding libpedoni_cuda.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=pedoni_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=PEDONI_CUDA_LIB_DIR");
}
