"""Drop-in boundary checks that need no GPU: the shared library loads, exports every symbol the headers
declare, the tables have the reference layout, and the host-only slots behave (errors, image I/O)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import dlimgedit_b200 as dl
from oracle import prepost as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built_library():
    if not os.path.exists(dl.LIB_PATH):
        from dlimgedit_b200 import _build
        _build.build()
    dl.load_library()


def test_exports_every_declared_symbol():
    declared = set()
    for name in os.listdir(os.path.join(ROOT, "include")):
        text = open(os.path.join(ROOT, "include", name)).read()
        declared |= set(re.findall(r"^DLIMG_B200_EXPORT [^;(]*?\b(\w+)\s*\(void\);", text, re.M))
    assert declared == set(dl.EXPORTED_SYMBOLS)
    out = subprocess.check_output(["nm", "-D", "--defined-only", dl.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert declared <= exported
    # nothing else leaks out of the library (the reference exports only dlimg_init for public-API users)
    assert {s for s in exported if not s.startswith("_")} == declared
    lib = ctypes.CDLL(dl.LIB_PATH)
    for sym in declared:
        assert getattr(lib, sym) is not None


def test_no_onnxruntime_or_vendor_library_dependency():
    out = subprocess.check_output(["ldd", dl.LIB_PATH], text=True)
    for banned in ("onnxruntime", "cublas", "cudnn", "torch", "nccl"):
        assert banned not in out.lower()


def test_table_layout_matches_reference_abi():
    # dlimgedit.h:44-68 -> 13 function pointers; dlimg_ImageView 24 bytes; dlimg_Options 16 bytes
    assert ctypes.sizeof(dl._Api) == 13 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(dl._ImageView) == 24 and dl._ImageView.pixels.offset == 16
    assert ctypes.sizeof(dl._Options) == 16 and dl._Options.model_directory.offset == 8
    api = dl.api()
    for name, _ in dl._Api._fields_:
        assert ctypes.cast(getattr(api, name), ctypes.c_void_p).value, name
    assert dl.ext().abi_version == 3


def test_backend_support_without_gpu():
    import torch
    assert dl.Environment.is_supported(dl.Backend.cpu) is False  # no CPU path in this engine
    if not torch.cuda.is_available():
        assert dl.Environment.is_supported(dl.Backend.gpu) is False


def test_environment_errors(tmp_path):
    with pytest.raises(dl.Exception, match="does not exist"):
        dl.Environment(dl.Options(dl.Backend.gpu, str(tmp_path / "nope")))
    f = tmp_path / "file"
    f.write_text("x")
    with pytest.raises(dl.Exception, match="is not a directory"):
        dl.Environment(dl.Options(dl.Backend.gpu, str(f)))
    with pytest.raises(dl.Exception, match="no CPU path"):
        dl.Environment(dl.Options(dl.Backend.cpu, str(tmp_path)))


def test_segment_objects_reports_unsupported():
    view = dl.ImageView(np.zeros((4, 4, 4), np.uint8)).to_c()
    out = np.zeros((4, 4), np.uint8)
    r = dl.api().segment_objects(ctypes.byref(view), out.ctypes.data, None)
    assert r == 1 and b"not supported" in dl.api().last_error()


def test_create_destroy_and_png_roundtrip(tmp_path):
    api = dl.api()
    p = api.create_image(16, 16, 4)
    assert p
    buf = (ctypes.c_uint8 * (16 * 16 * 4)).from_address(p)
    img = np.frombuffer(buf, np.uint8).reshape(16, 16, 4)
    for i in range(256):  # reference test/test_image.cpp:26-49
        img[i // 16, i % 16] = [255, i, 0, 255]
    view = dl._ImageView(16, 16, 4, 64, p)
    path = str(tmp_path / "save.png").encode()
    assert api.save_image(ctypes.byref(view), path) == 0
    ext = (ctypes.c_int * 2)()
    ch = ctypes.c_int()
    px = ctypes.c_void_p()
    assert api.load_image(path, ext, ctypes.byref(ch), ctypes.byref(px)) == 0
    assert (ext[0], ext[1], ch.value) == (16, 16, 4)
    got = np.frombuffer((ctypes.c_uint8 * 1024).from_address(px.value), np.uint8).reshape(16, 16, 4)
    assert (got == img).all()
    api.destroy_image(px)
    # a PNG written by another encoder (compressed, filtered) decodes too
    from PIL import Image
    rng = np.random.default_rng(0)
    ref = rng.integers(0, 255, (33, 21, 3), dtype=np.uint8)
    ref[5:20] = ref[4]  # compressible rows
    Image.fromarray(ref).save(tmp_path / "pil.png")
    assert api.load_image(str(tmp_path / "pil.png").encode(), ext, ctypes.byref(ch), ctypes.byref(px)) == 0
    assert (ext[0], ext[1], ch.value) == (21, 33, 3)
    got = np.frombuffer((ctypes.c_uint8 * ref.size).from_address(px.value), np.uint8).reshape(ref.shape)
    assert (got == ref).all()
    api.destroy_image(px)
    api.destroy_image(p)
    # unsupported save format and missing file report through last_error (image.cpp:26-29, 14-17)
    bad = dl._ImageView(16, 16, 5, 64, None)
    assert api.save_image(ctypes.byref(bad), path) == 1 and b"Unsupported channel order [5]" in api.last_error()
    assert api.load_image(b"/nonexistent.png", ext, ctypes.byref(ch), ctypes.byref(px)) == 1
    assert b"Failed to load image" in api.last_error()


@pytest.mark.parametrize("in_size,out_size", [(8, 4), (13, 18), (19, 26), (19, 10), (512, 1024), (1800, 1024),
                                              (1200, 683), (3840, 1024), (2160, 576), (1025, 1024), (100, 1024)])
def test_resize_plan_matches_oracle(in_size, out_size):
    """The engine's host-side resampling weights are bit-identical to the oracle's stb restatement."""
    first_o, w_o, n_o = P.resize_weights(in_size, out_size, 48)
    first = np.zeros(out_size, np.int32)
    w = np.zeros((out_size, 48), np.float32)
    n = dl.debug().resize_plan(in_size, out_size, 48, first.ctypes.data_as(dl.c_i32p), w.ctypes.data_as(dl.c_f32p))
    assert n > 0
    # the two tables may start a row at a different (zero-weight) tap: compare as dense rows over the input axis
    lo = min(int(first.min()), int(first_o.min()))
    def dense(f, ww):
        d = np.zeros((out_size, in_size + 64 - lo), np.float32)
        for o in range(out_size):
            d[o, f[o] - lo: f[o] - lo + 48] = ww[o]
        return d
    assert np.array_equal(dense(first, w).view(np.uint32), dense(first_o, w_o).view(np.uint32))


def test_srgb_tables_match_oracle():
    dec = np.zeros(256, np.float32)
    thr = np.zeros(256, np.float32)
    dl.debug().srgb_tables(dec.ctypes.data_as(dl.c_f32p), thr.ctypes.data_as(dl.c_f32p))
    assert np.array_equal(dec, P.srgb_decode_table())
    L = P.lib()
    for i in range(1, 256):  # threshold i is the first float that encodes to i
        t = thr[i]
        assert L.ref_linear_to_srgb8(float(t)) == i
        assert L.ref_linear_to_srgb8(float(np.nextafter(t, np.float32(-1)))) == i - 1


def test_load_image_reads_jpeg(tmp_path):
    """reference image.cpp:11-23 loads JPEG through stb_image; the library's own decoder restates stb's integer IDCT,
    hv_2 chroma upsampling and fixed-point YCbCr -> RGB.  Checked against libjpeg-turbo (PIL): the two decoders may
    differ by a few LSB (different IDCT rounding / upsampling rounding), not more."""
    from PIL import Image
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "truck.jpg")
    a = dl.api()
    ext = (ctypes.c_int * 2)()
    ch = ctypes.c_int()
    px = ctypes.c_void_p()
    assert a.load_image(path.encode(), ext, ctypes.byref(ch), ctypes.byref(px)) == 0, a.last_error()
    assert (ext[0], ext[1], ch.value) == (1800, 1200, 3)
    got = np.ctypeslib.as_array(ctypes.cast(px, ctypes.POINTER(ctypes.c_uint8)), shape=(1200, 1800, 3)).copy()
    a.destroy_image(px)
    ref = np.asarray(Image.open(path).convert("RGB"))
    d = np.abs(got.astype(int) - ref.astype(int))
    assert d.max() <= 4 and d.mean() < 0.1
    # 4:4:4, 4:2:2, greyscale and restart intervals, written by PIL from a synthetic picture
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:97, 0:131]
    pic = np.stack([127 + 100 * np.sin(xx / 9.0), 127 + 100 * np.cos(yy / 7.0), (xx + yy) % 256], -1)
    pic = np.clip(pic + rng.normal(0, 4, pic.shape), 0, 255).astype(np.uint8)
    for name, kw, mode in (("444.jpg", dict(subsampling=0), "RGB"), ("422.jpg", dict(subsampling=1), "RGB"),
                           ("420.jpg", dict(subsampling=2), "RGB"), ("grey.jpg", {}, "L")):
        f = str(tmp_path / name)
        Image.fromarray(pic if mode == "RGB" else pic[..., 0]).save(f, quality=92, **kw)
        assert a.load_image(f.encode(), ext, ctypes.byref(ch), ctypes.byref(px)) == 0, (name, a.last_error())
        n = ch.value
        got = np.ctypeslib.as_array(ctypes.cast(px, ctypes.POINTER(ctypes.c_uint8)), shape=(97, 131, n)).copy()
        a.destroy_image(px)
        ref = np.asarray(Image.open(f).convert(mode)).reshape(97, 131, n)
        d = np.abs(got.astype(int) - ref.astype(int))
        assert n == (3 if mode == "RGB" else 1) and d.max() <= 6 and d.mean() < 0.6, (name, d.max(), d.mean())
    # progressive files (spectral selection + successive approximation, one component per AC scan) hold the same
    # coefficients as their sequential twins, so they must decode to exactly the same pixels
    for name, kw, src in (("444", dict(subsampling=0), pic), ("422", dict(subsampling=1), pic), ("420", dict(subsampling=2), pic),
                          ("grey", {}, pic[..., 0])):
        for q in (35, 92):
            fb, fp = str(tmp_path / f"{name}_{q}_b.jpg"), str(tmp_path / f"{name}_{q}_p.jpg")
            Image.fromarray(src).save(fb, quality=q, **kw)
            Image.fromarray(src).save(fp, quality=q, progressive=True, **kw)
            assert open(fp, "rb").read().count(b"\xff\xc2") >= 1  # really SOF2
            base, prog = dl.Image.load(fb), dl.Image.load(fp)
            assert np.array_equal(base.pixels, prog.pixels), (name, q)
    rst = str(tmp_path / "restart.jpg")
    try:
        Image.fromarray(pic).save(rst, quality=80, progressive=True, restart_marker_blocks=3)
        has_rst = b"\xff\xdd" in open(rst, "rb").read()
    except TypeError:
        has_rst = False
    if has_rst:
        plain = str(tmp_path / "norestart.jpg")
        Image.fromarray(pic).save(plain, quality=80, progressive=True)
        assert np.array_equal(dl.Image.load(rst).pixels, dl.Image.load(plain).pixels)
    lossless = bytearray(open(str(tmp_path / "444.jpg"), "rb").read())
    lossless[lossless.index(b"\xff\xc0") + 1] = 0xC9  # pretend: arithmetic-coded frame
    (tmp_path / "arith.jpg").write_bytes(bytes(lossless))
    assert a.load_image(str(tmp_path / "arith.jpg").encode(), ext, ctypes.byref(ch), ctypes.byref(px)) == 1
    assert b"arithmetic" in a.last_error()


def test_load_image_reads_bmp_and_tga(tmp_path):
    """dlimgedit.hpp:59: "Supported formats are PNG, JPEG, BMP, TGA" (stb_image behind image.cpp:11-23).  Files written by
    PIL; decoded pixels must equal the source exactly (these formats are lossless)."""
    from PIL import Image
    rng = np.random.default_rng(5)
    pic = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    pic[5:15] = pic[4]
    pic[:, 10:30] = pic[:, 9:10]  # runs for the RLE forms
    rgba = np.dstack([pic, rng.integers(1, 256, (37, 53), dtype=np.uint8)])
    pal = Image.fromarray(pic).quantize(64)
    pal_rgb = np.asarray(pal.convert("RGB"))
    grey = pic[..., 0]
    cases = [("24.bmp", Image.fromarray(pic), {}, pic), ("32.bmp", Image.fromarray(rgba), {}, rgba),
             ("p8.bmp", pal, {}, pal_rgb), ("l8.bmp", Image.fromarray(grey), {}, np.dstack([grey] * 3)),
             ("1.bmp", Image.fromarray(grey > 127), {}, np.dstack([(grey > 127).astype(np.uint8) * 255] * 3)),
             ("24.tga", Image.fromarray(pic), {}, pic), ("24r.tga", Image.fromarray(pic), dict(compression="tga_rle"), pic),
             ("32.tga", Image.fromarray(rgba), {}, rgba), ("32r.tga", Image.fromarray(rgba), dict(compression="tga_rle"), rgba),
             ("l.tga", Image.fromarray(grey), {}, grey), ("lr.tga", Image.fromarray(grey), dict(compression="tga_rle"), grey),
             ("p.tga", pal, {}, pal_rgb), ("pr.tga", pal, dict(compression="tga_rle"), pal_rgb),
             ("top.tga", Image.fromarray(pic), dict(orientation=1), pic)]
    for name, img, kw, want in cases:
        f = tmp_path / name
        img.save(f, **kw)
        got = dl.Image.load(f)
        assert got.pixels.shape == want.shape, (name, got.pixels.shape)
        assert np.array_equal(got.pixels, want), name
    # 16-bit 5-5-5 BMP: bit patterns are replicated to 8 bits; a 32-bit file whose fourth byte is zero everywhere is opaque
    import struct
    w, h = 5, 3
    v = rng.integers(0, 1 << 15, (h, w), dtype=np.uint16)
    rows = b"".join(v[h - 1 - y].astype("<u2").tobytes() + b"\0" * ((4 - (2 * w) % 4) % 4) for y in range(h))
    hdr = b"BM" + struct.pack("<IHHI", 54 + len(rows), 0, 0, 54) + struct.pack("<IiiHHIIiiII", 40, w, h, 1, 16, 0, len(rows), 0, 0, 0, 0)
    (tmp_path / "555.bmp").write_bytes(hdr + rows)
    got = dl.Image.load(tmp_path / "555.bmp").pixels
    def rep(x):
        return (x << 3) | (x >> 2)
    want = np.dstack([rep((v >> 10) & 31), rep((v >> 5) & 31), rep(v & 31)]).astype(np.uint8)
    assert np.array_equal(got, want)
    bgr0 = np.dstack([pic[:h, :w, ::-1], np.zeros((h, w), np.uint8)])
    rows = b"".join(bgr0[h - 1 - y].tobytes() for y in range(h))
    hdr = b"BM" + struct.pack("<IHHI", 54 + len(rows), 0, 0, 54) + struct.pack("<IiiHHIIiiII", 40, w, -h, 1, 32, 0, len(rows), 0, 0, 0, 0)
    (tmp_path / "x32.bmp").write_bytes(hdr + b"".join(bgr0[y].tobytes() for y in range(h)))  # negative height: top-down
    got = dl.Image.load(tmp_path / "x32.bmp").pixels
    assert got.shape == (h, w, 4) and np.array_equal(got[..., :3], pic[:h, :w]) and (got[..., 3] == 255).all()
    a = dl.api()
    ext = (ctypes.c_int * 2)()
    ch = ctypes.c_int()
    px = ctypes.c_void_p()
    (tmp_path / "cut.bmp").write_bytes(open(tmp_path / "24.bmp", "rb").read()[:200])
    assert a.load_image(str(tmp_path / "cut.bmp").encode(), ext, ctypes.byref(ch), ctypes.byref(px)) == 1 and b"truncated" in a.last_error()
    (tmp_path / "cut.tga").write_bytes(open(tmp_path / "24r.tga", "rb").read()[:100])
    assert a.load_image(str(tmp_path / "cut.tga").encode(), ext, ctypes.byref(ch), ctypes.byref(px)) == 1 and b"truncated" in a.last_error()


def test_load_image_rejects_hostile_png_headers(tmp_path):
    """a 2^31 x 2^31 IHDR used to wrap the size arithmetic (heap overflow); short IHDR chunks were read past."""
    import struct
    import zlib

    def png(w, h, ihdr_len=13, payload=b"\x00" * 64):
        def chunk(t, d):
            return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
        ihdr = (struct.pack(">II", w, h) + bytes([8, 6, 0, 0, 0]))[:ihdr_len]
        return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(payload)) + chunk(b"IEND", b"")

    a = dl.api()
    ext = (ctypes.c_int * 2)()
    ch = ctypes.c_int()
    px = ctypes.c_void_p()
    for name, blob in (("huge.png", png(1 << 31, 1 << 31)), ("wide.png", png((1 << 24) + 1, 1)), ("zero.png", png(0, 5)),
                       ("short_ihdr.png", png(4, 4, ihdr_len=9)), ("bomb.png", png(2, 2, payload=b"\x00" * (1 << 20)))):
        f = tmp_path / name
        f.write_bytes(blob)
        r = a.load_image(str(f).encode(), ext, ctypes.byref(ch), ctypes.byref(px))
        if name == "bomb.png":  # more data than the image needs: refused before it is expanded
            assert r == 1 and b"larger than the image" in a.last_error()
        else:
            assert r == 1, name
    ok = tmp_path / "ok.png"
    ok.write_bytes(png(2, 2, payload=b"\x00" + bytes(8) + b"\x00" + bytes(8)))
    assert a.load_image(str(ok).encode(), ext, ctypes.byref(ch), ctypes.byref(px)) == 0, a.last_error()
    a.destroy_image(px)
    pnm = tmp_path / "overflow.ppm"
    pnm.write_bytes(b"P6 99999999999 99999999999 255\n" + bytes(12))
    assert a.load_image(str(pnm).encode(), ext, ctypes.byref(ch), ctypes.byref(px)) == 1


def test_image_class_owns_library_pixels(tmp_path):
    """dlimg::Image mirror (dlimgedit.hpp:48-82): pixels come from create_image / load_image and go back through
    destroy_image when the Image and its numpy views are gone.  Without an environment the storage is plain memory."""
    st = (ctypes.c_uint64 * 5)()
    dl.debug().image_pool_stats(st)
    plain0 = st[4]
    img = dl.Image(dl.Extent(320, 200), dl.Channels.rgb)
    assert img.pixels.shape == (200, 320, 3) and img.size() == 320 * 200 * 3
    img.pixels[:] = np.arange(320 * 3, dtype=np.uint8).reshape(1, 320, 3)
    dl.Image.save(img, tmp_path / "img.png")
    back = dl.Image.load(tmp_path / "img.png")
    assert (back.extent().width, back.extent().height, back.channels()) == (320, 200, dl.Channels.rgb)
    assert np.array_equal(back.pixels, img.pixels)
    rows = img.pixels[5:7]
    del img  # a view keeps the pixels alive
    assert rows[0, 1, 2] == 5
    mask = dl.Image(dl.Extent(7, 5), dl.Channels.mask)
    assert mask.pixels.shape == (5, 7)
    dl.debug().image_pool_stats(st)
    assert st[4] >= plain0 + 3 and st[0] == 0  # nothing is page-locked while no GPU environment exists
    assert dl.api().create_image(0, 4, 4) is None
    dl.api().destroy_image(None)


def _raw_load(path):
    a = dl.api()
    e = (ctypes.c_int * 2)()
    ch = ctypes.c_int()
    px = ctypes.c_void_p()
    assert a.load_image(str(path).encode(), e, ctypes.byref(ch), ctypes.byref(px)) == 0, a.last_error()
    arr = np.ctypeslib.as_array(ctypes.cast(px, ctypes.POINTER(ctypes.c_uint8)), shape=(e[1], e[0], ch.value)).copy()
    a.destroy_image(px)
    return arr


def test_load_image_reads_every_png_flavour(tmp_path):
    """stb_image (behind the reference's load_image) reads every PNG colour type and bit depth, palettes, tRNS and Adam7
    interlacing; 16-bit samples keep their high byte, sub-byte grey is scaled to 0..255, tRNS adds an alpha channel."""
    import struct
    import warnings
    import zlib
    from PIL import Image
    rng = np.random.default_rng(1)
    pic = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    pic[5:15] = pic[4]
    rgba = np.dstack([pic, rng.integers(0, 256, (37, 53), dtype=np.uint8)])
    grey = pic[..., 0]

    def check(name, img, want, **kw):
        img.save(tmp_path / name, **kw)
        got = _raw_load(tmp_path / name)
        want = want.reshape(want.shape[0], want.shape[1], -1)
        assert got.shape == want.shape and np.array_equal(got, want), name

    for colours, bits in ((37, 8), (13, 4), (4, 2), (2, 1)):  # palettes at every index width
        pal = Image.fromarray(pic).quantize(colours)
        check(f"p{bits}.png", pal, np.asarray(pal.convert("RGB")), bits=bits)
    palt = Image.fromarray(pic).quantize(20)
    palt.save(tmp_path / "pt.png", transparency=3)          # palette + tRNS -> 4 channels
    assert np.array_equal(_raw_load(tmp_path / "pt.png"), np.asarray(Image.open(tmp_path / "pt.png").convert("RGBA")))
    one = Image.fromarray(grey > 127)
    check("1.png", one, np.asarray(one).astype(np.uint8) * 255)  # 1-bit grey is scaled to 0 / 255
    check("l.png", Image.fromarray(grey), grey)
    check("la.png", Image.fromarray(np.dstack([grey, pic[..., 1]]), "LA"), np.dstack([grey, pic[..., 1]]))  # 2 channels, as stb reports
    with pytest.raises(dl.Exception, match="2 channels"):
        dl.Image.load(tmp_path / "la.png")                  # ... which the façade's Channels enum cannot name
    check("rgb.png", Image.fromarray(pic), pic)
    check("rgba.png", Image.fromarray(rgba), rgba)
    g16 = rng.integers(0, 65536, (37, 53), dtype=np.uint16)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        check("g16.png", Image.fromarray(g16, "I;16"), (g16 >> 8).astype(np.uint8))  # 16-bit: the high byte
    Image.fromarray(grey).save(tmp_path / "lk.png", transparency=int(grey[0, 0]))     # grey + colour key -> grey + alpha
    assert np.array_equal(_raw_load(tmp_path / "lk.png"), np.dstack([grey, np.where(grey == grey[0, 0], 0, 255).astype(np.uint8)]))
    Image.fromarray(pic).save(tmp_path / "rk.png", transparency=tuple(int(v) for v in pic[0, 0]))
    alpha = np.where((pic == pic[0, 0]).all(-1), 0, 255).astype(np.uint8)
    assert np.array_equal(_raw_load(tmp_path / "rk.png"), np.dstack([pic, alpha]))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)

    def adam7(arr, colour):  # PIL writes no interlaced files: the seven passes by hand, filter type 0
        h, w = arr.shape[:2]
        s = arr.reshape(h, w, -1)
        raw = b""
        for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
            sub = s[y0::dy, x0::dx]
            if sub.size:
                raw += b"".join(b"\x00" + row.tobytes() for row in sub)
        ihdr = struct.pack(">IIBBBBB", w, h, 8, colour, 0, 0, 1)
        return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")

    for name, arr, colour in (("i_rgb.png", pic, 2), ("i_small.png", pic[:3, :5].copy(), 2), ("i_rgba.png", rgba, 6), ("i_grey.png", grey[:9, :2].copy(), 0)):
        (tmp_path / name).write_bytes(adam7(arr, colour))
        assert np.array_equal(_raw_load(tmp_path / name), arr.reshape(arr.shape[0], arr.shape[1], -1)), name
    a = dl.api()
    ext = (ctypes.c_int * 2)()
    ch = ctypes.c_int()
    px = ctypes.c_void_p()
    bad = bytearray(adam7(pic, 2))
    bad[24] = 3  # bit depth 3 does not exist
    (tmp_path / "bad_depth.png").write_bytes(bytes(bad))
    assert a.load_image(str(tmp_path / "bad_depth.png").encode(), ext, ctypes.byref(ch), ctypes.byref(px)) == 1


def test_save_image_writes_compressed_png(tmp_path):
    """save_image (reference image.cpp:25-35 -> stbi_write_png): filtered scanlines in a fixed-Huffman deflate stream.
    Any PNG reader must get the pixels back; a mask shrinks by two orders of magnitude; noise falls back to stored blocks."""
    from PIL import Image
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:600, 0:900]
    mask = (((xx - 450) ** 2 + (yy - 300) ** 2 < 200 ** 2) * 255).astype(np.uint8)
    photo = np.clip(np.stack([127 + 100 * np.sin(xx / 37.0), 127 + 100 * np.cos(yy / 23.0), (xx + yy) % 256], -1)
                    + rng.normal(0, 3, (600, 900, 3)), 0, 255).astype(np.uint8)
    cases = [("mask.png", mask, dl.Channels.mask), ("photo.png", photo, dl.Channels.rgb),
             ("rgba.png", np.dstack([photo[:100, :150], mask[:100, :150]]), dl.Channels.rgba),
             ("noise.png", rng.integers(0, 256, (97, 131, 3), dtype=np.uint8), dl.Channels.rgb),
             ("one.png", np.array([[9]], np.uint8), dl.Channels.mask), ("flat.png", np.full((5, 300, 3), 7, np.uint8), dl.Channels.rgb)]
    sizes = {}
    for name, arr, ch in cases:
        img = dl.Image(dl.Extent(arr.shape[1], arr.shape[0]), ch)
        img.pixels[:] = arr
        dl.Image.save(img, tmp_path / name)
        assert np.array_equal(np.asarray(Image.open(tmp_path / name)), arr), name      # libpng reads it
        assert np.array_equal(dl.Image.load(tmp_path / name).pixels, arr), name        # and so does the library
        sizes[name] = (tmp_path / name).stat().st_size
    assert sizes["mask.png"] < mask.size // 50
    assert sizes["photo.png"] < photo.size * 0.9
    assert sizes["noise.png"] < 97 * 131 * 3 + 300       # stored blocks: no worse than raw + framing
    assert sizes["flat.png"] < 200
