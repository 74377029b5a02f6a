#!/usr/bin/env python
"""TEST INFRASTRUCTURE — builds libibmfoam_b200_emu.so: the WHOLE C-ABI library (ibmfoam_b200/csrc/*.cu, unchanged) for
the CPU emulation of tools/cuda_emu, so that the GPU parity tests can be run on a machine without a GPU:

    python tools/cuda_emu/build_emu.py
    IBM_TEST_EMU_LIB=tools/cuda_emu/_build/libibmfoam_b200_emu.so python -m pytest tests -m gpu -k "not fused" ...

The only source transformation is syntactic: `kernel<<<grid, block, smem, stream>>>(args)` becomes a call of the
emulator's launch function, `cudaLaunchCooperativeKernel((void*)kernel, ...)` one that still knows the kernel's type, and `extern __shared__ T name[]` a pointer to the emulated dynamic shared memory; everything
else is handled by the stand-in headers (cuda_runtime.h, cub/cub.cuh, cooperative_groups.h, cuda_pipeline.h).
It checks LOGIC.  It is slow (every CUDA thread is a fiber), says nothing about performance or about the hardware memory
model, and is never loaded by the product: only tests/conftest.py knows IBM_TEST_EMU_LIB (it re-points the ctypes loader
for the test session); ibmfoam_b200.capi itself has no switch and the real library refuses to run without a CUDA device."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "ibmfoam_b200", "csrc")
OUT = os.path.join(HERE, "_build")
SRC_OUT = os.path.join(OUT, "src")
LIB = os.path.join(OUT, "libibmfoam_b200_emu.so")


def _match(text, i, open_ch, close_ch):
    """index just past the bracket that closes the one at text[i]"""
    depth = 0
    while i < len(text):
        c = text[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced brackets")


def _split_top(s):
    out, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    out.append("".join(cur).strip())
    return out


def transform(text):
    out, pos = [], 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            out.append(text[pos:])
            break
        # kernel expression: identifier, optionally followed by <template args>
        j = k
        while j > 0 and text[j - 1].isspace():
            j -= 1
        if text[j - 1] == ">":
            depth, j = 0, j - 1
            while True:
                if text[j] == ">":
                    depth += 1
                elif text[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        while j > 0 and (text[j - 1].isalnum() or text[j - 1] in "_:"):
            j -= 1
        kern = text[j:k].strip()
        e = text.find(">>>", k)
        cfg = _split_top(text[k + 3:e])
        while len(cfg) < 4:
            cfg.append("0")
        a0 = e + 3
        while text[a0].isspace():
            a0 += 1
        assert text[a0] == "(", text[k - 40:k + 80]
        a1 = _match(text, a0, "(", ")")
        args = text[a0 + 1:a1 - 1].strip()
        out.append(text[pos:j])
        out.append(f"EMU_LAUNCH(({kern}), {cfg[0]}, {cfg[1]}, {cfg[2]}, {cfg[3]}" + (f", {args}" if args else "") + ")")
        pos = a1
    text = "".join(out)
    text = re.sub(r"cudaLaunchCooperativeKernel\(\s*\(void\s*\*\)\s*(\w+)\s*,", r"::emu::coop_launch_erased(\1,", text)
    text = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\s*\[\s*\]\s*;", r"\1* \2 = (\1*)::emu::dyn_smem();", text)
    return text


def build(verbose=False, asan=False):
    """asan=True: a second library (…_asan.so) instrumented with AddressSanitizer - device memory is host memory here, so
    an out-of-bounds global or static-shared access of a kernel is reported like compute-sanitizer's memcheck would
    (run the tests with LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0)."""
    global SRC_OUT, LIB
    if asan:
        SRC_OUT = os.path.join(OUT, "src_asan")
        LIB = os.path.join(OUT, "libibmfoam_b200_emu_asan.so")
    os.makedirs(SRC_OUT, exist_ok=True)
    objs = []
    units = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    flags = ["-x", "c++", "-std=c++17", "-O1", "-g", *os.environ.get("IBM_EMU_DEFS", "").split(), "-fPIC", "-ffp-contract=off", "-w", "-I" + HERE, "-I" + CSRC, "-I" + os.path.join(ROOT, "include")]
    if asan:
        flags += ["-fsanitize=address", "-fno-omit-frame-pointer"]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in ("emu.hpp", "cuda_runtime.h", "cuda_pipeline.h",
                                                                                                 "cooperative_groups.h", "build_emu.py")]
    deps.append(os.path.join(HERE, "cub", "cub.cuh"))
    newest = max(os.path.getmtime(d) for d in deps)
    for u in units:
        src = os.path.join(SRC_OUT, u.replace(".cu", ".emu.cpp"))
        obj = src.replace(".emu.cpp", ".o")
        objs.append(obj)
        if os.path.exists(obj) and os.path.getmtime(obj) > newest:
            continue
        with open(os.path.join(CSRC, u)) as f:
            text = transform(f.read())
        with open(src, "w") as f:
            # csrc headers are included by their own names; relative includes of the original location keep working through -I
            f.write(f'#line 1 "{os.path.join(CSRC, u)}"\n' + text)
        r = subprocess.run(["g++", *flags, "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stderr[-6000:])
            raise RuntimeError(f"emulation build failed on {u}")
        if verbose:
            print("compiled", u)
    r = subprocess.run(["g++", "-shared", "-o", LIB, *objs, "-ldl", "-lpthread"] + (["-fsanitize=address"] if asan else []),
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr[-6000:])
        raise RuntimeError("emulation link failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, asan="--asan" in sys.argv))
