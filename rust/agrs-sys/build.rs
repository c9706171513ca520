// The shared libraries are built by `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a) into
// autograd.rs_b200/; point AGRS_LIB_DIR at that directory.
fn main() {
    let dir = std::env::var("AGRS_LIB_DIR").expect("set AGRS_LIB_DIR to the directory holding libagrs_b200.so and libagrs_engine.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=AGRS_LIB_DIR");
}
