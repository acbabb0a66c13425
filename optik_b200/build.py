"""Build recipe: optik_b200/csrc -> optik_b200/lib/liboptik_b200.so (sm_100a only, in-tree).

    python -m optik_b200.build [--force] [--verbose]

-fmad=false is part of the arithmetic spec (see csrc/dmath.cuh): fused multiply-adds are explicit fma() calls.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "liboptik_b200.so")
SRCS = ["solve_kernel.cu", "solve_t1_kernel.cu", "eval_kernel.cu", "diffik_kernel.cu", "peak_kernel.cu", "exchange_kernel.cu", "robot.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def sources():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "optik_b200.h"))
    return deps


def build(force=False, verbose=False, defines=(), out=OUT):
    """`defines` / `out`: experiment builds (tools/exp_variants.py) of the same sources under other -D switches."""
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in sources()):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
           "-shared", "-Xcompiler", "-fPIC,-O2,-fvisibility=default", "-o", out]
    cmd += ["-D" + d for d in defines]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SRCS]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
