"""Builds libdlimgedit.so (C++ host + sm_100a CUDA kernels) in-tree with nvcc.

    python -m dlimgedit_b200._build [--force] [--verbose] [--dev]

--dev compiles the development switches and their alternative kernels in (-DDLIMG_B200_DEV, csrc/common.hpp); a change of the
flag rebuilds everything.

The library is linked against the static CUDA runtime only: no cuBLAS / cuDNN / onnxruntime.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB_PATH = os.path.join(HERE, "libdlimgedit.so")  # reference library name, src/CMakeLists.txt:20-23

SOURCES = [
    "api.cu",
    "engine.cu",
    "model.cu",
    "weights.cpp",
    "image_io.cpp",
    "image_pool.cpp",
    "profiler.cpp",
    "kernels/gemm.cu",
    "kernels/encoder_kernels.cu",
    "kernels/window_attention.cu",
    "kernels/patch_embed.cu",
    "kernels/mbconv_tail.cu",
    "kernels/local_conv.cu",
    "kernels/decoder_kernels.cu",
    "kernels/decoder_tokens.cu",
    "kernels/t2i_attention.cu",
    "kernels/prepost_kernels.cu",
]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
COMMON = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unknown-pragmas",
    "-DDLIMG_B200_BUILD",
    "--expt-relaxed-constexpr",
]


def _headers():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".hpp", ".cuh", ".h"))]
    inc = os.path.join(os.path.dirname(HERE), "include")
    out += [os.path.join(inc, f) for f in os.listdir(inc)]
    return out


def _compile(src: str, obj: str, verbose: bool, dev: bool = False):
    cmd = [NVCC] + COMMON + (["-DDLIMG_B200_DEV"] if dev else []) + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return r.stderr


def build(force: bool = False, verbose: bool = False, dev: bool = False) -> str:
    os.makedirs(os.path.join(OBJ_DIR, "kernels"), exist_ok=True)
    flavour = os.path.join(OBJ_DIR, "flavour")
    want = "dev" if dev else "release"
    if not os.path.exists(flavour) or open(flavour).read() != want:
        force = True
    newest_header = max(os.path.getmtime(h) for h in _headers())
    jobs = []
    objs = []
    for rel in SOURCES:
        src = os.path.join(CSRC, rel)
        obj = os.path.join(OBJ_DIR, rel.rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header)
        if stale:
            jobs.append((src, obj))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(lambda j: _compile(j[0], j[1], verbose, dev), jobs):
                if verbose and log:
                    print(log)
    if jobs or not os.path.exists(LIB_PATH):
        cmd = [NVCC, "-shared", "-o", LIB_PATH] + objs + ["-cudart", "static", "-Xlinker", "-soname,libdlimgedit.so.1"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    soname = LIB_PATH + ".1"  # SOVERSION 1 like the reference (src/CMakeLists.txt:20-23); a copy, so it survives any sync
    if not os.path.exists(soname) or os.path.getmtime(soname) < os.path.getmtime(LIB_PATH):
        import shutil
        shutil.copy2(LIB_PATH, soname)
    with open(flavour, "w") as f:
        f.write(want)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, dev="--dev" in sys.argv))
