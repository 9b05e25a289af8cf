// Links the prebuilt libhfb200.so (built by `python -c 'import __graft_entry__ as g; g.build()'`).
fn main() {
    let dir = std::env::var("HFB200_LIB_DIR").unwrap_or_else(|_| "../../".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=hfb200");
    println!("cargo:rerun-if-env-changed=HFB200_LIB_DIR");
}
