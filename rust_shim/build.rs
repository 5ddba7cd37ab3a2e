// Links libp2g.so (built by acvm-backend-plonky2_b200/csrc/Makefile).  P2G_LIB_DIR = the directory that holds it.
fn main() {
    let dir = std::env::var("P2G_LIB_DIR").unwrap_or_else(|_| "../acvm-backend-plonky2_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=p2g");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=P2G_LIB_DIR");
}
