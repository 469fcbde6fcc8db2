// Links libmzb200.so (built by `python __graft_entry__.py` / nvcc).  cudarc-free: the library
// links the CUDA runtime statically, so nothing else is needed at link time.
fn main() {
    let dir = std::env::var("MZB200_LIB_DIR").unwrap_or_else(|_| "../../simd-minimizers_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=mzb200");
    println!("cargo:rerun-if-env-changed=MZB200_LIB_DIR");
}
