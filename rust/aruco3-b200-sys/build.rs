// Links against the prebuilt CUDA library; point ARUCO3_B200_LIB_DIR at the directory holding libaruco3_b200.so.
fn main() {
    if let Ok(dir) = std::env::var("ARUCO3_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=aruco3_b200");
    println!("cargo:rerun-if-env-changed=ARUCO3_B200_LIB_DIR");
}
