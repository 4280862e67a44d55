//! Locates libspf_b200.so.  The library is built by `make` at the root of the spf_b200 repository
//! (nvcc -gencode arch=compute_100a,code=sm_100a); point SPF_B200_LIB_DIR at the directory that holds it
//! (default: ../../spf_b200 relative to this crate, i.e. the in-tree build).
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("SPF_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../spf_b200")
    });
    let dir = dir.canonicalize().unwrap_or(dir);
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=spf_b200");
    // the shared object carries its own CUDA runtime (static cudart); only the driver is needed at run time
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=SPF_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=build.rs");
}
