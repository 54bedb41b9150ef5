// Links the prebuilt libreve_cuda.so (make -C reve_b200/csrc).  REVE_CUDA_LIB_DIR overrides the path.
fn main() {
    let dir = std::env::var("REVE_CUDA_LIB_DIR").unwrap_or_else(|_| "../reve_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=reve_cuda");
    println!("cargo:rerun-if-env-changed=REVE_CUDA_LIB_DIR");
}
