"""In-tree build of libvmap_b200.so (nvcc, sm_100a only).  No JIT cache: the .so sits next to
this file so it travels with the repository snapshot to the GPU box."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvmap_b200.so")
SOURCES = ["vmap.cu"]


def deps():
    """Every file the library is compiled from: all of csrc/ plus the public header."""
    import glob
    out = []
    for pat in ("*.cu", "*.cuh", "*.inl", "*.h"):
        out += glob.glob(os.path.join(CSRC, pat))
    out.append(os.path.join(HERE, "..", "include", "vmap", "vmap.h"))
    out.append(os.path.abspath(__file__))
    return out


NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # bit-exact parity: no FMA contraction anywhere (the kernels also use explicit *_rn
    # intrinsics on every order-sensitive operation)
    "-fmad=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared",
]
LINK_FLAGS = ["-ldl"]


def nvcc_path():
    p = shutil.which("nvcc")
    if p:
        return p
    p = "/usr/local/cuda/bin/nvcc"
    if os.path.exists(p):
        return p
    raise RuntimeError("nvcc not found: libvmap_b200.so cannot be built (there is no CPU fallback)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in deps())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + LINK_FLAGS
    if verbose:
        cmd += ["-Xptxas", "-v"]
    subprocess.check_call(cmd, cwd=HERE)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
