// Links the in-tree C-ABI library (probly_search_b200/_lib/libprobly_b200.so).
fn main() {
    let dir = std::env::var("PROBLY_B200_LIB_DIR").unwrap_or_else(|_| "../probly_search_b200/_lib".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=probly_b200");
}
