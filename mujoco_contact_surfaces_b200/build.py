"""In-tree build of libhcs_b200.so (hand-written CUDA for sm_100a + C++ host code, C ABI in include/hcs.h).

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.
`-fmad=false` / `-ffp-contract=off`: the clipping kernels must reproduce the fp64 operation order of
the restated Drake path bit for bit (DESIGN.md "Floating-point contract").
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# tuning sweeps: HCS_VARIANT=name HCS_NVCC_DEFS="-DBP_CTAS_PER_SM=5 ..." builds variants/libhcs_b200.<name>.so,
# which HCS_LIB=<path> makes engine.py load instead of the in-tree library
VARIANT = os.environ.get("HCS_VARIANT", "")
OUT = os.path.join(HERE, "variants", "libhcs_b200.%s.so" % VARIANT) if VARIANT else os.path.join(HERE, "libhcs_b200.so")
OBJ = os.path.join(CSRC, "_build" + ("_" + VARIANT if VARIANT else ""))

SOURCES = ["engine.cu", "kernels_build.cu", "kernels_step.cu", "kernels_broadphase.cu", "kernels_tactile.cu", "kernels_lbvh.cu", "kernels_meshgen.cu",
           "mesh_host.cpp", "multi.cpp"]

NVCC = os.environ.get("HCS_NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-ccbin", "g++", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2,-pthread", "-Xptxas", "-v",
] + os.environ.get("HCS_NVCC_DEFS", "").split()


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + [
        os.path.join(HERE, "..", "include", "hcs.h"), os.path.abspath(__file__)]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile(src):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    path = os.path.join(CSRC, src)
    if not _stale(obj, [path] + _deps()):
        return obj, ""
    cmd = [NVCC] + NVCC_FLAGS + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(_compile, SOURCES))
    objs = [o for o, _ in results]
    log = "\n".join(l for _, l in results if l)
    if log:
        with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
            f.write(log)
        if verbose:
            print(log)
    if _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "g++", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return OUT


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
