"""Builds tools/pfslam_run: the headless driver with the reference's main.cpp call sequence
(src/main.cpp:175-237) over the reference's own kernel.h symbols (libpfslam_kernelh.so).

Test / demonstration tooling, not product: it is compiled against the reference's headers and links the
reference's scene parser (scene.o, utilities.o -- compiled from /root/reference by oracle/Makefile), so
it is only built where those exist; the binary travels to the GPU box.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "gpu-icp-slam_b200")
RUN_BIN = os.path.join(HERE, "pfslam_run")


def build_pfslam_run(ref="/root/reference"):
    src = os.path.join(HERE, "pfslam_run.cpp")
    objs = [os.path.join(ROOT, "oracle", "_ref", o) for o in ("scene.o", "utilities.o")]
    libs = [os.path.join(PKG, "libpfslam_kernelh.so"), os.path.join(PKG, "libpfslam.so")]
    if not os.path.isdir(os.path.join(ref, "src")) or not all(os.path.exists(p) for p in objs + libs):
        return None
    if os.path.exists(RUN_BIN) and os.path.getmtime(RUN_BIN) >= max(os.path.getmtime(p) for p in [src] + objs + libs):
        return RUN_BIN
    cmd = [os.environ.get("CXX", "g++"), "-std=c++14", "-O2", "-w",
           "-DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP", "-I/usr/local/cuda/include",
           "-I" + os.path.join(ref, "external", "include"), "-I" + os.path.join(ref, "src"),
           "-o", RUN_BIN, src] + objs + ["-L" + PKG, "-lpfslam_kernelh", "-lpfslam",
           "-Wl,-rpath,$ORIGIN/../gpu-icp-slam_b200", "-L/usr/local/cuda/lib64", "-lcudart"]
    subprocess.run(cmd, check=True)
    return RUN_BIN


if __name__ == "__main__":
    print(build_pfslam_run())
