// Links against the prebuilt C-ABI library (python -m flo_b200.build -> flo_b200/libflo_b200.so).
// FLO_B200_LIB_DIR overrides the search path.
fn main() {
    let dir = std::env::var("FLO_B200_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{}/../../flo_b200", manifest)
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=flo_b200");
    println!("cargo:rerun-if-env-changed=FLO_B200_LIB_DIR");
}
