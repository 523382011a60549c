"""Builds ``librabacus_b200.so`` in-tree with nvcc for sm_100a.

Flags: ``-fmad=false`` keeps plain ``a*b+c`` unfused like the reference's
gfortran build (fusion is explicit ``fma()`` in the hot loops); ``-lineinfo``
for ncu source pages; static cudart so the library has no torch dependency.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "rb_api.cu")
OUT = os.path.join(HERE, "librabacus_b200.so")
DEPS = [os.path.join(HERE, "csrc", f) for f in sorted(os.listdir(os.path.join(HERE, "csrc")))
        if f.endswith((".cu", ".cuh", ".h"))]
DEPS.append(os.path.join(os.path.dirname(HERE), "include", "rabacus_b200.h"))

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
              "-fmad=false", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC,-ffp-contract=off"]


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


OUT_CHECKED = os.path.join(HERE, "librabacus_b200_checked.so")


def build_checked(verbose=False):
    """The checked build: -DRB_DEBUG_CHECKS turns on the device-side assertions (RB_ASSERT)
    and the guard words behind every plan buffer.  Not built by default."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    flags = [f for f in NVCC_FLAGS if f != "-O3"] + ["-O2", "-DRB_DEBUG_CHECKS"]
    cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT_CHECKED, SRC]
    print("[rabacus_b200] " + " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return OUT_CHECKED


def build(force=False, verbose=False):
    if up_to_date() and not force:
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build %s" % OUT)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    print("[rabacus_b200] " + " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    if "--checked" in sys.argv:
        build_checked(verbose="-v" in sys.argv)
    else:
        build(force="--force" in sys.argv, verbose="-v" in sys.argv)
