"""
TEST INFRASTRUCTURE ONLY.  Builds the CPU oracle (oracle/fa_oracle.c) into oracle/_build/libfa_oracle.so.

The reference itself cannot be compiled (`oracle/_ref`): it is browser JavaScript and this image has no
JS engine (node / deno / bun / qjs absent) -- see DESIGN.md "Oracle".  So there is no oracle/_ref binary; instead
oracle/minijs executes the reference's own minified modules in the build container (outputs committed as
tests/golden/ref_js.json), and where `node` exists oracle/run_reference_modules.js does the same.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libfa_oracle.so")
SRC = os.path.join(HERE, "fa_oracle.c")
DEPS = [SRC] + [os.path.join(ROOT, "include", h) for h in ("fa_b200.h", "fa_jsmath.h", "fa_tables.h", "fa_curves.h")]


def _cpu_has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            return " fma " in f.read()
    except OSError:
        return False


def ensure_built(force: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS):
        return LIB
    if not os.path.exists(SRC):  # prebuilt .so shipped without sources is fine
        if os.path.exists(LIB):
            return LIB
        raise FileNotFoundError(SRC)
    cmd = ["gcc", "-std=gnu11", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-fopenmp",
           "-fvisibility=hidden", "-I", os.path.join(ROOT, "include"), "-o", LIB + ".tmp", SRC, "-lm"]
    if _cpu_has_fma():
        cmd.insert(2, "-mfma")  # fmaf() -> one vfmadd instruction; identical results to libm's fmaf, just faster
    subprocess.check_call(cmd)
    os.replace(LIB + ".tmp", LIB)
    return LIB


def build_gate_enum(force: bool = False) -> str:
    """oracle/gate/gate_enum.c: the exhaustive walk of the noise gate's y -> v map (tests/golden/make_gate_golden.py)."""
    src, exe = os.path.join(HERE, "gate", "gate_enum.c"), os.path.join(BUILD, "gate_enum")
    os.makedirs(BUILD, exist_ok=True)
    if not force and os.path.exists(exe) and os.path.getmtime(exe) >= max(os.path.getmtime(src), os.path.getmtime(DEPS[2])):
        return exe
    subprocess.check_call(["gcc", "-std=gnu11", "-O2", "-fopenmp", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                           "-o", exe + ".tmp", src, "-lm"])
    os.replace(exe + ".tmp", exe)
    return exe


if __name__ == "__main__":
    print(ensure_built(force="--force" in sys.argv))
