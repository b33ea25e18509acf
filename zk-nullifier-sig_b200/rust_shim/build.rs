// Link against the in-tree C-ABI library: PLUME_B200_LIB_DIR=<repo>/zk-nullifier-sig_b200
fn main() {
    let dir = std::env::var("PLUME_B200_LIB_DIR").unwrap_or_else(|_| "..".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=plume_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
}
