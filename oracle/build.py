"""Builds the oracle's C restatement (oracle/c/prepost_ref.c) into oracle/_build/libprepost_ref.so.

Test infrastructure only.  There is no `oracle/_ref`: the reference's own path cannot be compiled from
its few sources -- it needs onnxruntime 1.20.1, Eigen, stb and fmt, all fetched by cmake at configure
time and absent offline -- so it is treated as unbuildable.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libprepost_ref.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "c", "prepost_ref.c")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    # -ffp-contract=off: keep the float operation order of the restated algorithm (no fused multiply-add)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, src, "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
