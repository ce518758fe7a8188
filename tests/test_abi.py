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
    assert dl.ext().abi_version == 2


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
