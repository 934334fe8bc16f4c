"""In-tree builds of the native libraries (no JIT cache: the .so files travel with the repo).

  csrc/libvgi.so        — CUDA kernels + C ABI (include/vgi.h), sm_100a only
  csrc/libvgi_synth.so  — host-side synthetic input producers (plain C, not on the hot path)
"""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIBVGI = os.path.join(CSRC, "libvgi.so")
LIBSYNTH = os.path.join(CSRC, "libvgi_synth.so")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

# Kernels whose results are quantised (occupancy bits, RGBA8 texels) are compiled without FMA
# contraction so they follow the same IEEE binary32 contract as the oracle (DESIGN.md "numerics").
CU_STRICT = ["vgi_build.cu", "vgi_svo.cu", "vgi_atlas.cu", "vgi_raster.cu"]
CU_FAST = ["vgi_trace.cu", "vgi_post.cu"]
CPP = ["vgi_api.cpp"]
HEADERS = ["vgi_internal.h", "vgi_device.cuh", os.path.join("..", "..", "include", "vgi.h")]

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off", "-ccbin", GXX]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def build_libvgi(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in CU_STRICT + CU_FAST + CPP]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS]
    if not force and not _stale(LIBVGI, deps):
        return LIBVGI
    objs = []
    for f in CU_STRICT + CU_FAST + CPP:
        src = os.path.join(CSRC, f)
        obj = os.path.join(CSRC, os.path.splitext(f)[0] + ".o")
        if force or _stale(obj, [src] + [os.path.join(CSRC, h) for h in HEADERS]):
            cmd = [NVCC] + ARCH + COMMON + ["-c", src, "-o", obj]
            if f in CU_STRICT:
                cmd += ["-fmad=false"]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            subprocess.check_call(cmd)
        objs.append(obj)
    subprocess.check_call([NVCC] + ARCH + ["-shared", "-o", LIBVGI] + objs + ["-ccbin", GXX, "-lcudart"])
    return LIBVGI


def build_synth(force=False):
    src = os.path.join(CSRC, "synth_raster.c")
    if not force and not _stale(LIBSYNTH, [src, os.path.join(_HERE, "..", "include", "vgi.h")]):
        return LIBSYNTH
    subprocess.check_call([GCC, "-O2", "-std=gnu11", "-fPIC", "-fopenmp", "-shared", "-o", LIBSYNTH, src, "-lm"])
    return LIBSYNTH


def build_all(force=False, verbose=False):
    build_synth(force)
    build_libvgi(force, verbose)


if __name__ == "__main__":
    import sys
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIBVGI, LIBSYNTH)
