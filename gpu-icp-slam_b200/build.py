"""Builds libpfslam.so (the CUDA engine + C ABI) in-tree with nvcc for sm_100a.

    python gpu-icp-slam_b200/build.py [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libpfslam.so")
SOURCES = [os.path.join(HERE, "csrc", "pfslam.cu")]
DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + [
    os.path.join(ROOT, "include", "pfslam.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.run(cmd, check=True)
    return LIB


COMPAT_LIB = os.path.join(HERE, "libpfslam_kernelh.so")


def build_kernel_h_compat(ref="/root/reference"):
    """The reference's kernel.h entry points over the C ABI (csrc/kernel_h_compat.cpp).  Needs the
    reference's headers, so it is only built where /root/reference exists; the .so travels."""
    if not os.path.isdir(os.path.join(ref, "src")):
        return None
    src = os.path.join(HERE, "csrc", "kernel_h_compat.cpp")
    if os.path.exists(COMPAT_LIB) and os.path.getmtime(COMPAT_LIB) >= max(os.path.getmtime(src), os.path.getmtime(LIB)):
        return COMPAT_LIB
    cmd = [os.environ.get("CXX", "g++"), "-std=c++14", "-O2", "-fPIC", "-shared", "-w",
           "-DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP", "-I/usr/local/cuda/include",
           "-I" + os.path.join(ref, "external", "include"), "-I" + os.path.join(ref, "src"),
           "-o", COMPAT_LIB, src, "-L" + HERE, "-lpfslam", "-Wl,-rpath,$ORIGIN",
           "-L/usr/local/cuda/lib64", "-lcudart"]
    subprocess.run(cmd, check=True)
    return COMPAT_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_kernel_h_compat())
