// Links libb200prover.so (built by `make -C renegade_b200/csrc`): point B200PROVER_LIB_DIR at the directory holding it.
fn main() {
    let dir = std::env::var("B200PROVER_LIB_DIR")
        .expect("set B200PROVER_LIB_DIR to the directory holding libb200prover.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=b200prover");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=B200PROVER_LIB_DIR");
}
