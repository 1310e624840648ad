// build.rs for the reference crate once it links libsbx: compiles the .cu sources with nvcc through `cc`
// (or links a prebuilt libsbx.so when SBX_LIB_DIR is set).  NOT compiled in this repository's image (no
// cargo/rustc); kept as the source a maintainer would add next to Cargo.toml.
fn main() {
    if let Ok(dir) = std::env::var("SBX_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=sbx");
        return;
    }
    let csrc = std::path::Path::new("sbx/soap_bubble_simulation_b200/csrc");
    cc::Build::new()
        .cuda(true)
        .flag("-gencode")
        .flag("arch=compute_100a,code=sm_100a")
        .flag("-std=c++17")
        .flag("-O3")
        .flag("-lineinfo")
        .files(["sbx_kernels.cu", "sbx_march_tma.cu", "sbx_wgsl_tma.cu", "sbx_shade_tma.cu", "sbx_resident.cu", "sbx_frame.cu",
                "sbx_api.cu", "sbx_comm.cu"].iter().map(|f| csrc.join(f)))  // SRCS of csrc/Makefile
        .compile("sbx");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=dl");
    println!("cargo:rustc-link-lib=z");  // sbx_export_frame_png
}
