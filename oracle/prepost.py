"""ORACLE (test infrastructure only) -- numpy/ctypes front-end over oracle/c/prepost_ref.c plus the
composed CPU reference pipeline for the segmentation hot path (reference src/segmentation.cpp:121-174).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

_lib = None

MASK, RGB, RGBA, BGRA, ARGB = 1, 3, 4, 5, 6  # Channels enum values (dlimgedit.hpp:29)


def bytes_per_pixel(channels: int) -> int:
    return 4 if channels > 4 else channels  # dlimgedit.impl.hpp:15


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB if os.path.exists(_build.LIB) else _build.build()
        L = ctypes.CDLL(path)
        u8p, f32p, i32p = (ctypes.POINTER(t) for t in (ctypes.c_uint8, ctypes.c_float, ctypes.c_int))
        L.ref_resize_longest_side.argtypes = [ctypes.c_int] * 3 + [i32p, i32p, f32p]
        L.ref_resize_longest_side.restype = ctypes.c_int
        L.ref_transform_coord.argtypes = [ctypes.c_int, ctypes.c_float]
        L.ref_transform_coord.restype = ctypes.c_int
        L.ref_create_image_tensor.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p]
        L.ref_write_mask_image.argtypes = [f32p] + [ctypes.c_int] * 5 + [u8p]
        L.ref_resize_srgb.argtypes = [u8p] + [ctypes.c_int] * 4 + [u8p, ctypes.c_int, ctypes.c_int]
        L.ref_resize_srgb.restype = ctypes.c_int
        L.ref_srgb_decode_table.argtypes = [f32p]
        L.ref_linear_to_srgb8.argtypes = [ctypes.c_float]
        L.ref_linear_to_srgb8.restype = ctypes.c_uint8
        L.ref_resize_weights.argtypes = [ctypes.c_int] * 3 + [i32p, f32p]
        L.ref_resize_weights.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def resize_longest_side(w: int, h: int, max_side: int = 1024):
    """-> (needs_resize, out_w, out_h, scale)"""
    ow, oh, sc = ctypes.c_int(), ctypes.c_int(), ctypes.c_float()
    r = lib().ref_resize_longest_side(w, h, max_side, ctypes.byref(ow), ctypes.byref(oh), ctypes.byref(sc))
    return bool(r), ow.value, oh.value, np.float32(sc.value)


def transform_coord(c: int, scale) -> int:
    return lib().ref_transform_coord(int(c), float(scale))


def create_image_tensor(pixels: np.ndarray, channels: int) -> np.ndarray:
    """pixels: packed uint8 (h, w, bpp) -> float32 (h, w, 3), values 0..255"""
    pixels = np.ascontiguousarray(pixels, dtype=np.uint8)
    h, w = pixels.shape[:2]
    out = np.empty((h, w, 3), np.float32)
    lib().ref_create_image_tensor(_p(pixels, ctypes.c_uint8), w, h, channels, _p(out, ctypes.c_float))
    return out


def write_mask_image(logits: np.ndarray, index: int, w: int, h: int) -> np.ndarray:
    """logits: float32 (1, n, th, tw) -> uint8 (h, w) of 0/255"""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    _, _, th, tw = logits.shape
    out = np.empty((h, w), np.uint8)
    lib().ref_write_mask_image(_p(logits, ctypes.c_float), th, tw, index, w, h, _p(out, ctypes.c_uint8))
    return out


def resize_srgb(pixels: np.ndarray, out_w: int, out_h: int, stride: int = 0) -> np.ndarray:
    """pixels uint8: either (h, w, c) packed, or a flat strided buffer with explicit (h, w, c) given by
    `pixels.shape` and byte `stride`."""
    h, w, c = pixels.shape
    if stride:
        buf = pixels  # caller passes a view into a strided buffer
        base = buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
    else:
        buf = np.ascontiguousarray(pixels, dtype=np.uint8)
        base = _p(buf, ctypes.c_uint8)
    out = np.empty((out_h, out_w, c), np.uint8)
    ok = lib().ref_resize_srgb(base, w, h, stride, c, _p(out, ctypes.c_uint8), out_w, out_h)
    assert ok
    return out


def srgb_decode_table() -> np.ndarray:
    t = np.empty(256, np.float32)
    lib().ref_srgb_decode_table(_p(t, ctypes.c_float))
    return t


def resize_weights(in_size: int, out_size: int, max_taps: int = 40):
    first = np.empty(out_size, np.int32)
    w = np.empty((out_size, max_taps), np.float32)
    n = lib().ref_resize_weights(in_size, out_size, max_taps, _p(first, ctypes.c_int), _p(w, ctypes.c_float))
    assert n >= 0
    return first, w, n


def prompt_tensors(point, region, scale):
    """reference src/segmentation.cpp:134-152: -> (coords (1,2,2) f32, labels (1,2) f32)"""
    coords = np.zeros((1, 2, 2), np.float32)
    labels = np.zeros((1, 2), np.float32)

    def put(i, x, y, lab):
        coords[0, i, 0] = float(transform_coord(x, scale))
        coords[0, i, 1] = float(transform_coord(y, scale))
        labels[0, i] = float(lab)

    assert (point is None) != (region is None)
    if point is not None:
        put(0, point[0], point[1], 1)
        put(1, 0, 0, -1)
    else:
        put(0, region[0], region[1], 2)
        put(1, region[2], region[3], 3)
    return coords, labels
