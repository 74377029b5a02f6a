"""In-tree nvcc build of libibmfoam_b200.so for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libibmfoam_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-Xptxas", "-v"]

# (source, extra flags).  The reference-order TU must not contract a*b+c into FMAs: its membership
# arithmetic has to match an x86-64 build of the reference bit for bit.
UNITS = [
    ("ibm_kernels_ref.cu", ["-fmad=false"]),
    ("ibm_kernels_v2.cu", []),
    ("ibm_kernels_v7.cu", []),
    ("ibm_kernels_alpha.cu", ["-fmad=false"]),
    ("ibm_kernels_util.cu", []),
    ("ibm_motion.cu", ["-fmad=false"]),
    ("ibm_comm.cu", []),
    ("ibm_capi.cu", []),
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "ibmfoam_b200.h"))
    objs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC, *ARCH, *COMMON, *extra, "-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log = os.path.join(LIBDIR, src.replace(".cu", ".ptxas.log"))
            with open(log, "w") as f:
                f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"nvcc failed on {src}")
            if verbose:
                sys.stderr.write(r.stderr)
    if force or _stale(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    # host-side C++ mirror of the reference class API: compile its driver against the C ABI (runs on the GPU box)
    host = os.path.join(HERE, "host")
    exe = os.path.join(LIBDIR, "test_host_mirror")
    srcs = [os.path.join(host, "test_host_mirror.cpp"), os.path.join(host, "particleCloud.H")]
    if force or _stale(exe, srcs + [LIB]):
        cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-o", exe, srcs[0], "-L" + LIBDIR, "-libmfoam_b200", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("host mirror driver failed to build")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
