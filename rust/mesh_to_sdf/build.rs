// Links libm2s.so. M2S_LIB_DIR points at the directory that holds it (mesh_to_sdf_b200/ in this repo).
fn main() {
    let dir = std::env::var("M2S_LIB_DIR").unwrap_or_else(|_| "../../mesh_to_sdf_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=m2s");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=M2S_LIB_DIR");
}
