// NOT BUILT IN THIS REPOSITORY'S IMAGE (no cargo / rustc).  Points the linker at libpcp_b200.so.
fn main() {
    let dir = std::env::var("PCP_B200_LIB_DIR").unwrap_or_else(|_| "../pcp_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=pcp_b200");
    println!("cargo:rerun-if-env-changed=PCP_B200_LIB_DIR");
}
