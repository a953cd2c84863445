// Link against the in-tree shared library: ZARU_B200_LIB_DIR = <repo>/zaru_b200 (where `python __graft_entry__.py` puts it).
fn main() {
    let dir = std::env::var("ZARU_B200_LIB_DIR").unwrap_or_else(|_| "../../zaru_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=zaru_b200");
    println!("cargo:rerun-if-env-changed=ZARU_B200_LIB_DIR");
}
