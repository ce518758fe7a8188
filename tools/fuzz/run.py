"""Mutation fuzzing of the image readers (load_image: PNG / JPEG / BMP / TGA / PNM) under AddressSanitizer + UBSan.

    python tools/fuzz/run.py [seed [rounds]]

Builds csrc/image_io.cpp + csrc/image_pool.cpp with g++ -fsanitize=address,undefined into /tmp (no GPU needed: without an
environment the pixel buffers are plain memory), writes seed files of every format with PIL, and feeds the reader mutated copies
(byte flips, truncation, header corruption, huge 32-bit fields, 0xFF / bit flips in the entropy-coded part).  Every file must either
decode or fail with an exception; any sanitizer report stops the run and names the file.
"""
import os
import struct
import subprocess
import sys
import zlib

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
WORK = "/tmp/dlimg_fuzz"
CSRC = os.path.join(ROOT, "dlimgedit_b200", "csrc")


def build():
    os.makedirs(WORK, exist_ok=True)
    exe = os.path.join(WORK, "fuzz")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-I", CSRC, "-I", "/usr/local/cuda/include", os.path.join(ROOT, "tools", "fuzz", "fuzz_image_io.cpp"),
                           os.path.join(CSRC, "image_io.cpp"), os.path.join(CSRC, "image_pool.cpp"),
                           "-L/usr/local/cuda/lib64", "-lcudart", "-o", exe])
    return exe


def seeds():
    d = os.path.join(WORK, "seed")
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(0)
    pic = rng.integers(0, 256, (45, 61, 3), dtype=np.uint8)
    pic[5:15] = pic[4]
    pic[:, 10:30] = pic[:, 9:10]
    rgba = np.dstack([pic, rng.integers(0, 256, (45, 61), dtype=np.uint8)])
    out = []

    def sv(name, img, **kw):
        img.save(os.path.join(d, name), **kw)
        out.append(os.path.join(d, name))

    sv("a.jpg", Image.fromarray(pic), quality=80)
    sv("b.jpg", Image.fromarray(pic), quality=80, progressive=True)
    sv("c.jpg", Image.fromarray(pic), quality=60, subsampling=2, progressive=True)
    sv("g.jpg", Image.fromarray(pic[..., 0]), progressive=True)
    sv("a.png", Image.fromarray(pic))
    sv("b.png", Image.fromarray(rgba))
    sv("c.png", Image.fromarray(pic).quantize(16), bits=4)
    sv("d.png", Image.fromarray(pic).quantize(20), transparency=3)
    sv("e.png", Image.fromarray(pic[..., 0] > 100))
    sv("a.bmp", Image.fromarray(pic))
    sv("b.bmp", Image.fromarray(rgba))
    sv("c.bmp", Image.fromarray(pic).quantize(16))
    sv("d.bmp", Image.fromarray(pic[..., 0] > 100))
    sv("a.tga", Image.fromarray(pic))
    sv("b.tga", Image.fromarray(rgba), compression="tga_rle")
    sv("c.tga", Image.fromarray(pic).quantize(16), compression="tga_rle")
    sv("d.tga", Image.fromarray(pic[..., 0]))

    def chunk(t, data):
        return struct.pack(">I", len(data)) + t + data + struct.pack(">I", zlib.crc32(t + data) & 0xffffffff)

    raw = b""
    for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
        sub = pic[y0::dy, x0::dx]
        if sub.size:
            raw += b"".join(b"\x00" + r.tobytes() for r in sub)
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 61, 45, 8, 2, 0, 0, 1)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")
    for name, blob in (("i.png", png), ("p.ppm", b"P6 61 45 255\n" + pic.tobytes())):
        with open(os.path.join(d, name), "wb") as f:
            f.write(blob)
        out.append(os.path.join(d, name))
    return out


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    exe = build()
    files0 = seeds()
    mut = os.path.join(WORK, "mut")
    os.makedirs(mut, exist_ok=True)
    rng = np.random.default_rng(seed)
    total = 0
    for rnd in range(rounds):
        files = []
        for s in files0:
            data = bytearray(open(s, "rb").read())
            for k in range(60):
                d = bytearray(data)
                mode = rng.integers(0, 5)
                if mode == 0:
                    for _ in range(int(rng.integers(1, 6))):
                        d[int(rng.integers(0, len(d)))] = int(rng.integers(0, 256))
                elif mode == 1:
                    d = d[:int(rng.integers(1, len(d)))]
                elif mode == 2:
                    for _ in range(int(rng.integers(1, 4))):
                        d[int(rng.integers(0, min(len(d), 64)))] = int(rng.integers(0, 256))
                elif mode == 3:
                    p = int(rng.integers(0, max(1, len(d) - 4)))
                    d[p:p + 4] = struct.pack(">I", int(rng.choice([0, 1, 0x7fffffff, 0xffffffff, 0x80000000, 65536, 1 << 24])))
                else:
                    for _ in range(int(rng.integers(1, 20))):
                        q = int(rng.integers(len(d) // 3, len(d)))
                        d[q] = int(rng.choice([0xFF, 0x00, d[q] ^ (1 << int(rng.integers(0, 8)))]))
                f = os.path.join(mut, f"{os.path.basename(s)}.{rnd}.{k}")
                with open(f, "wb") as fh:
                    fh.write(bytes(d))
                files.append(f)
        r = subprocess.run([exe] + files, capture_output=True, text=True, timeout=1800)
        total += len(files)
        print(rnd, r.returncode, r.stdout.strip(), flush=True)
        if r.returncode:
            print(r.stderr[-3000:])
            for f in files:
                if subprocess.run([exe, f], capture_output=True, timeout=120).returncode:
                    print("culprit:", f)
                    break
            sys.exit(1)
        for f in files:
            os.remove(f)
    print("files", total, "clean")


if __name__ == "__main__":
    main()
