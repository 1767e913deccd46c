"""Builds libopb.so (hand-written sm_100a CUDA behind the C ABI of include/opb.h) in-tree.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "opb_api.cu")
LIB = os.path.join(HERE, os.environ.get("OPB_LIB_NAME", "libopb.so"))
DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + [
    os.path.join(os.path.dirname(HERE), "include", "opb.h")]


def needs_build():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build_native(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC", "-Xptxas", "-v" if verbose else "-O3", "-o", LIB, SRC, "-ldl"] + os.environ.get("OPB_NVCC_FLAGS", "").split()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed building libopb.so")
    if verbose:
        print(r.stdout)
    return LIB


if __name__ == "__main__":
    build_native(force=True, verbose="-v" in sys.argv)
    print("built", LIB)
