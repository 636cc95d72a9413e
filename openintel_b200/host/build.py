"""Builds the C++ host layer: libopenintel_host.so (tokenizer + index builder shim, CPU only) and
host_demo (links libopenintel_gpu.so).  Run by __graft_entry__.build()."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
CUDA_LIB = "/usr/local/cuda/lib64"


def _newer(out, srcs):
    return not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs)


def build():
    hdr = [os.path.join(HERE, "openintel_host.hpp"), os.path.join(HERE, "openintel_store.hpp"), os.path.join(PKG, "..", "include", "openintel_gpu.h")]
    so = os.path.join(HERE, "libopenintel_host.so")
    src = os.path.join(HERE, "host_capi.cpp")
    if _newer(so, [src] + hdr):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", src, "-o", so, "-ldl"])
    demo = os.path.join(HERE, "host_demo")
    src = os.path.join(HERE, "host_demo.cpp")
    gpu_so = os.path.join(PKG, "libopenintel_gpu.so")
    if os.path.exists(gpu_so) and _newer(demo, [src, gpu_so] + hdr):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", src, "-o", demo, "-L" + PKG, "-lopenintel_gpu",
                               "-L" + CUDA_LIB, "-ldl", "-Wl,-rpath," + PKG, "-Wl,-rpath," + CUDA_LIB, "-Wl,-rpath-link," + CUDA_LIB])
    return so, demo


if __name__ in ("__main__", "__build__"):
    print(build())
