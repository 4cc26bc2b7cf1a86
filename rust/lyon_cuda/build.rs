// Links liblyon_cuda.so (built by `python -c "import __graft_entry__ as g; g.build()"` -> lyon_b200/liblyon_cuda.so).
fn main() {
    let dir = std::env::var("LYON_CUDA_LIB_DIR").unwrap_or_else(|_| "../../lyon_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=lyon_cuda");
    println!("cargo:rerun-if-env-changed=LYON_CUDA_LIB_DIR");
}
