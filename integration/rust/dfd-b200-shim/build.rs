// Link against the C-ABI library built by `python -c "import __graft_entry__ as g; g.build()"` (or the nvcc recipe in
// datafusion_distributed_b200/build.py).  DFD_B200_LIB_DIR = directory holding libdfd_b200.so.
fn main() {
    let dir = std::env::var("DFD_B200_LIB_DIR").unwrap_or_else(|_| "../../../datafusion_distributed_b200/_lib".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=dfd_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=DFD_B200_LIB_DIR");
}
