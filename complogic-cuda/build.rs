// Link against the prebuilt libcomplogic_cuda.so (host C++ + embedded sm_100a cubin; it dlopens
// libcuda.so.1 itself, so no CUDA toolkit is needed to build or link the crate).
fn main() {
  let dir = std::env::var("COMPLOGIC_CUDA_LIB_DIR").unwrap_or_else(|_| "../complogic_b200".into());
  println!("cargo:rustc-link-search=native={dir}");
  println!("cargo:rustc-link-lib=dylib=complogic_cuda");
  println!("cargo:rerun-if-env-changed=COMPLOGIC_CUDA_LIB_DIR");
}
