//! Links `libspectre_b200.so` (built by `python -m spectre_b200.build` in the spectre-b200 repo) into the forked
//! halo2_proofs crate. `SPECTRE_B200_LIB_DIR` must point at the directory holding the library; an rpath is embedded so
//! the `spectre-prover` binary finds it at run time without LD_LIBRARY_PATH.
use std::{env, path::PathBuf};

fn main() {
    println!("cargo:rerun-if-env-changed=SPECTRE_B200_LIB_DIR");
    if env::var_os("CARGO_FEATURE_B200").is_none() {
        return; // CPU-only build of the fork: nothing to link
    }
    let dir = PathBuf::from(
        env::var_os("SPECTRE_B200_LIB_DIR")
            .expect("set SPECTRE_B200_LIB_DIR to the directory that contains libspectre_b200.so"),
    );
    assert!(
        dir.join("libspectre_b200.so").exists(),
        "{} does not contain libspectre_b200.so (run `python -m spectre_b200.build`)",
        dir.display()
    );
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=spectre_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
}
