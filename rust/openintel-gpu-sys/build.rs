// build.rs — compiles the CUDA sources for sm_100a with nvcc and links the shared library.
// Same recipe as openintel_b200/_build.py (the one that is exercised in the source repo).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("OPENINTEL_CSRC").unwrap_or_else(|_| "../../openintel_b200/csrc".into()));
    let srcs = ["api.cu", "cosine_scan.cu", "cosine_gemm.cu", "bm25.cu", "rrf.cu", "comm.cu", "lexicon.cu", "synth.cu"];
    let mut objs = Vec::new();
    for s in srcs {
        let o = out.join(format!("{s}.o"));
        let mut cmd = Command::new("nvcc");
        cmd.args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                  "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]);
        if s == "bm25.cu" || s == "rrf.cu" {
            cmd.arg("-fmad=false"); // docs/SPEC.md §3/§4: one IEEE f32 operation per source-level operation
        }
        let ok = cmd.arg("-c").arg(csrc.join(s)).arg("-o").arg(&o).status().expect("nvcc not found").success();
        assert!(ok, "nvcc failed on {s}");
        objs.push(o);
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    let lib = out.join("libopenintel_gpu.so");
    let ok = Command::new("nvcc").arg("-shared").arg("-o").arg(&lib).args(&objs)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl", "-lpthread"])
        .status().expect("nvcc not found").success();
    assert!(ok, "link failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=openintel_gpu");
}
