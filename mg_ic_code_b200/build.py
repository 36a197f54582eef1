"""Builds libmgic_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

-fmad=false: no FMA contraction, so every kernel keeps the reference Fortran's operation order
 and is bit-comparable with a non-contracting CPU evaluation.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SO = os.path.join(LIBDIR, "libmgic_b200.so")
SOURCES = ["alloc.cu", "kernels.cu", "gsrb_fused.cu", "restrict_tma.cu", "bottom.cu", "bottom_brick.cu", "bottom_dsmem.cu", "bottom_cbrick.cu", "source.cu", "grids.cu", "capi.cu", "chf_abi.cu", "comm.cu"]
HEADERS = [os.path.join(CSRC, "mgic_internal.h"), os.path.join(CSRC, "mgic_device.cuh"), os.path.join(CSRC, "arena.h"), os.path.join(CSRC, "tma.cuh"), os.path.join(ROOT, "include", "mgic.h"),
           os.path.join(ROOT, "include", "mgic_chf.h"), os.path.join(ROOT, "include", "mgic_comm.h")]


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + [h for h in HEADERS if os.path.exists(h)] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every translation unit for sm_100a (objects in parallel), link the shared library in-tree."""
    if not force and not stale():
        return SO
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "_obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17", "-diag-suppress", "177", "-Xcompiler", "-fPIC",
             "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]
    if verbose:
        flags.insert(0, "-Xptxas=-v")

    def one(src):
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        cmd = [nvcc] + flags + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(one, sources()))
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO] + objs)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(SO)
