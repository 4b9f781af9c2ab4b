"""Build libpsgd_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m psgd_torch_b200.build [--force] [-v]

Every .cu of csrc/ is compiled to an object file under csrc/_obj/ (in parallel, skipped when the object is newer than all
sources/headers) and the objects are linked into the shared library.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libpsgd_b200.so")
SOURCES = ["api.cu", "gemm_simt.cu", "gemm_tc.cu", "lra.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-DPSGD_NO_FAST_MATH"]


# headers each translation unit includes (directly or through common.cuh); the public header reaches all of them
DEPS = {"api.cu": ["common.cuh", "kron_kernels.cuh", "kron_geom.cuh", "bounds.cuh", "tc_ptx.cuh"], "gemm_simt.cu": ["common.cuh"],
        "gemm_tc.cu": ["common.cuh", "tc_ptx.cuh"],
        "lra.cu": ["common.cuh", "lra_mma.cuh", "lra_tc.cuh", "tc_ptx.cuh"]}


def _headers(src):
    deps = [os.path.join(CSRC, f) for f in DEPS.get(src, [f for f in os.listdir(CSRC) if f.endswith(".cuh")])]
    deps.append(os.path.join(HERE, "..", "include", "psgd_b200.h"))
    return [d for d in deps if os.path.exists(d)]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    jobs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        if force or _newer(obj, [src] + _headers(s)):
            jobs.append((src, obj))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if not jobs and not _newer(LIB, objs):
        return LIB

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        return subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=max(1, len(jobs))) as ex:
        results = list(ex.map(compile_one, jobs))
    for r in results:
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed building libpsgd_b200.so")
        if verbose:
            print(r.stdout + r.stderr)
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libpsgd_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
