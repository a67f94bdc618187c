// Link against the C-ABI library built by `python -m quokkasim_b200.build` (quokkasim_b200/libquokka_b200.so).
fn main() {
    let dir = std::env::var("QUOKKA_B200_LIB_DIR").unwrap_or_else(|_| "../quokkasim_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=quokka_b200");
    println!("cargo:rerun-if-env-changed=QUOKKA_B200_LIB_DIR");
}
