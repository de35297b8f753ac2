// Links the prebuilt C-ABI library; set CHALKYDRI_B200_LIB_DIR to the directory holding libchalkydri_b200.so.
fn main() {
    if let Ok(dir) = std::env::var("CHALKYDRI_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=chalkydri_b200");
    println!("cargo:rerun-if-env-changed=CHALKYDRI_B200_LIB_DIR");
}
