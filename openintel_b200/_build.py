"""Build recipe of libopenintel_gpu.so: plain nvcc for sm_100a, in-tree, no JIT cache.

`python -m openintel_b200._build` or `__graft_entry__.build()`.  The .so is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
import glob
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libopenintel_gpu.so")
OBJ = os.path.join(HERE, "csrc", "_obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unknown-pragmas", "--expt-relaxed-constexpr",
    "-ccbin", "/usr/bin/g++",
]
# bm25.cu must keep one IEEE op per source op (SPEC §3): no FMA contraction there
PER_FILE = {"bm25.cu": ["-fmad=false"], "rrf.cu": ["-fmad=false"]}


def _nvcc():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: cannot build libopenintel_gpu.so")
    return p


def _stamp(src, flags):
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())
    deps = [src] + sorted(glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh"))
                          + glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    for d in deps:
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    objs, rebuilt = [], False
    procs = []
    for s in srcs:
        name = os.path.basename(s)
        flags = NVCC_FLAGS + PER_FILE.get(name, [])
        o = os.path.join(OBJ, name + ".o")
        st = o + ".stamp"
        want = _stamp(s, flags)
        objs.append(o)
        if not force and os.path.exists(o) and os.path.exists(st) and open(st).read() == want:
            continue
        cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((name, st, want, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, st, want, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on " + name)
        if verbose:
            sys.stderr.write(out)
        open(st, "w").write(want)
        rebuilt = True
    if rebuilt or not os.path.exists(OUT):
        cmd = [nvcc, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++",
                                                     "-lcudart", "-ldl", "-lpthread"]
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
