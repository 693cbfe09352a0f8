// build.rs — link against libkc_b200.so built by `make -C uniprot_kmer_based_clustering_b200/csrc`.
// KC_B200_LIB_DIR overrides the in-tree location.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("KC_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap())
            .join("..")
            .join("uniprot_kmer_based_clustering_b200")
            .join("lib")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=kc_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=KC_B200_LIB_DIR");
}
