"""Mutation fuzzing of the weight-container reader (csrc/weights.cpp, WeightFile::load) under AddressSanitizer + UBSan.

    python tools/fuzz/run_weights.py [seed [files]]

Every mutated container must either load or fail with an exception; header sizes are checked against the file size before
anything is allocated (a corrupt tensor count used to ask for tens of gigabytes)."""
import os
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dlimgedit_b200 import weights_io  # noqa: E402

WORK = "/tmp/dlimg_fuzz_weights"
CSRC = os.path.join(ROOT, "dlimgedit_b200", "csrc")


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n_files = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    os.makedirs(os.path.join(WORK, "mut"), exist_ok=True)
    exe = os.path.join(WORK, "fuzz")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-I", CSRC,
                           "-I", "/usr/local/cuda/include", os.path.join(ROOT, "tools", "fuzz", "fuzz_weights.cpp"),
                           os.path.join(CSRC, "weights.cpp"), "-o", exe])
    rng = np.random.default_rng(seed)
    tensors = {"a.weight": rng.normal(size=(4, 3, 2)).astype(np.float32), "b.bias": rng.normal(size=(7,)).astype(np.float32),
               "c": rng.normal(size=(1,)).astype(np.float32)}
    seed_file = os.path.join(WORK, "seed.bin")
    weights_io.save(seed_file, tensors)
    data = open(seed_file, "rb").read()
    big = [0, 1, 2 ** 63, 2 ** 64 - 1, 2 ** 40, 2 ** 32, 2 ** 62, 2 ** 61 + 5]
    small = [0, 1, 0x7fffffff, 0xffffffff, 0x80000000, 65536, 1 << 24]
    files = []
    for k in range(n_files):
        d = bytearray(data)
        mode = int(rng.integers(0, 4))
        if mode == 0:
            for _ in range(int(rng.integers(1, 5))):
                d[int(rng.integers(0, len(d)))] = int(rng.integers(0, 256))
        elif mode == 1:
            d = d[:int(rng.integers(1, len(d)))]
        elif mode == 2:
            p = int(rng.integers(8, min(len(d) - 8, 120)))
            d[p:p + 4] = struct.pack("<I", small[int(rng.integers(0, len(small)))])
        else:
            p = int(rng.integers(8, min(len(d) - 8, 120)))
            d[p:p + 8] = struct.pack("<Q", big[int(rng.integers(0, len(big)))])
        f = os.path.join(WORK, "mut", f"{k}.bin")
        with open(f, "wb") as fh:
            fh.write(bytes(d))
        files.append(f)
    r = subprocess.run([exe, seed_file] + files, capture_output=True, text=True, timeout=1800)
    print(r.returncode, r.stdout.strip())
    if r.returncode:
        print(r.stderr[-3000:])
        sys.exit(1)
    print("files", len(files), "clean")


if __name__ == "__main__":
    main()
