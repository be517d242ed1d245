// Links liblrge_b200.so (include/lrge_b200.h).  LRGE_B200_LIB_DIR = directory holding the shared library
// (lrge_b200/csrc in this repository after `make`).
fn main() {
    let dir = std::env::var("LRGE_B200_LIB_DIR").expect("set LRGE_B200_LIB_DIR to the directory holding liblrge_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=lrge_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=LRGE_B200_LIB_DIR");
}
