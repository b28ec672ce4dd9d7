// build.rs — replaces the reference's build.rs (which compiles src/{cpu|gpu}/*.{cpp|cu} into libmatrix.a,
// build.rs:27-92): builds libcg_b200.so for sm_100a with the repository's own Makefile (nvcc), and links it.
//
//   CG_B200_DIR   path of the computational-graph_b200 checkout (default: the parent of this crate)
//
// There is no CPU fallback: without nvcc the build fails, and without a GPU the first call into the library reports
// "no CUDA device available".
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let root = env::var("CG_B200_DIR")
        .map(PathBuf::from)
        .unwrap_or_else(|_| PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join(".."));
    let csrc = root.join("computational-graph_b200").join("csrc");
    let status = Command::new("make")
        .arg("-C")
        .arg(&csrc)
        .arg("libcg_b200.so")
        .status()
        .expect("could not run make (is nvcc on PATH? this backend is CUDA-only)");
    assert!(status.success(), "building libcg_b200.so failed");
    println!("cargo:rustc-link-search=native={}", csrc.display());
    println!("cargo:rustc-link-lib=dylib=cg_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
