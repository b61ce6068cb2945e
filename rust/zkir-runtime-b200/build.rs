// Link against the in-tree CUDA library.  ZKIR_B200_LIB_DIR = directory that holds libzkir_b200.so.
fn main() {
    let dir = std::env::var("ZKIR_B200_LIB_DIR").unwrap_or_else(|_| "../../zkir_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=zkir_b200");
    println!("cargo:rerun-if-env-changed=ZKIR_B200_LIB_DIR");
}
