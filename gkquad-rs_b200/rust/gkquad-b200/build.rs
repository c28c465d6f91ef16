// UNCOMPILED (no Rust toolchain in the build image).
fn main() {
    let dir = std::env::var("GKQUAD_B200_LIB_DIR").unwrap_or_else(|_| "../../lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=gkquad_b200");
}
