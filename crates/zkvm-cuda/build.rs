// Builds the CUDA engine with nvcc through the `cc` crate (north_star: "a thin extern "C" FFI crate built with
// cc/nvcc") and links it.  LATTICE_AJTAI_SRC points at this repository's latticeum_b200/csrc and include/.
fn main() {
    let src = std::env::var("LATTICE_AJTAI_SRC").unwrap_or_else(|_| "../../latticeum_b200/csrc".into());
    println!("cargo:rerun-if-changed={src}");
    cc::Build::new()
        .cuda(true)
        .cudart("static")
        .flag("-gencode")
        .flag("arch=compute_100a,code=sm_100a")
        .flag("-O3")
        .flag("-lineinfo")
        .flag("-std=c++17")
        .file(format!("{src}/engine.cu"))
        .file(format!("{src}/ring_kernels.cu"))
        .file(format!("{src}/mac_kernels.cu"))
        .file(format!("{src}/ntt_pow2.cu"))
        .compile("lattice_ajtai");
    println!("cargo:rustc-link-lib=stdc++");
}
