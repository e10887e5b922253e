// build.rs — compile the CUDA sources with nvcc and link the static/shared library
fn main() {
    let out = std::path::PathBuf::from(std::env::var("OUT_DIR").unwrap());
    let csrc = "b200/csrc";                       // = halo2_aes_b200/csrc of this repository
    let mut objs = vec![];
    for f in ["api.cu", "witness.cu", "lookup.cu", "ntt.cu", "quotient.cu", "msm.cu"] {
        let o = out.join(f).with_extension("o");
        let ok = std::process::Command::new("nvcc")
            .args(["-O3", "-std=c++17", "-lineinfo",
                   "-gencode", "arch=compute_100a,code=sm_100a",
                   "-Xcompiler", "-fPIC", "-c"])
            .arg(format!("{csrc}/{f}")).arg("-o").arg(&o)
            .status().expect("nvcc").success();
        assert!(ok, "nvcc failed on {f}");
        objs.push(o);
    }
    let lib = out.join("libh2aes.a");
    assert!(std::process::Command::new("ar").arg("crs").arg(&lib).args(&objs)
        .status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=h2aes");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=static=cudart_static");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rerun-if-changed={csrc}");
}
