"""Build recipe for the oracle's C restatement -> oracle/_build/liboptik_oracle.so.

The reference itself (Rust + un-vendored NLopt) cannot be compiled in this
image (no cargo/rustc), so there is no oracle/_ref; see DESIGN.md.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "liboptik_oracle.so")
SRCS = [os.path.join(HERE, f) for f in ("optik_oracle.c", "solver_twin.c", "ref_loop.c", "diffik_oracle.c")]


def build(force: bool = False) -> str:
    srcs = [s for s in SRCS if os.path.exists(s)]
    hdrs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".h")]
    if not force and os.path.exists(OUT) and all(
        os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs + hdrs
    ):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # -ffp-contract=off: every fused multiply-add in the solver twin is an
    # explicit fma() so that it is bit-identical to the CUDA kernel (which is
    # compiled with -fmad=false).  -mfma only makes fma() a single instruction.
    cmd = ["gcc", "-O3", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
           "-pthread", "-o", OUT] + srcs + ["-lm"]
    import platform
    if platform.machine() in ("x86_64", "AMD64"):
        try:
            flags = open("/proc/cpuinfo").read()
            if " fma " in flags:
                cmd.insert(1, "-mfma")
        except OSError:
            pass
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
