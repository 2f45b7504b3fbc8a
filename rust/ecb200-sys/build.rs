// Links the prebuilt C-ABI library.  ECB200_LIB_DIR = directory holding libecb200.so (built by
// `python rustcrypto-elliptic-curves_b200/build.py`, nvcc for sm_100a); there is no CPU fallback to build instead.
fn main() {
    if let Ok(dir) = std::env::var("ECB200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=ecb200");
    println!("cargo:rerun-if-env-changed=ECB200_LIB_DIR");
}
