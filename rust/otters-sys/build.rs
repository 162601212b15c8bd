// Links against the prebuilt libotters_b200.so (built by `make -C otters_b200/csrc` with nvcc for sm_100a).
fn main() {
    let dir = std::env::var("OTTERS_B200_LIB_DIR").unwrap_or_else(|_| "../../otters_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=otters_b200");
    println!("cargo:rerun-if-env-changed=OTTERS_B200_LIB_DIR");
}
