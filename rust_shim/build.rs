// Links the C ABI library built by `make -C pcp_b200/csrc` (nvcc, sm_100a).
fn main() {
    let dir = std::env::var("PCP_B200_LIB_DIR").unwrap_or_else(|_| "../pcp_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=pcp_b200");
}
