"""dlimgedit_b200 -- Python mirror of dlimgedit's public C++ façade for the segmentation hot path
(reference src/include/dlimgedit/dlimgedit.hpp:16-193), bound with ctypes to the C ABI of the
B200-native engine (include/dlimg_b200.h -> dlimgedit_b200/libdlimgedit.so).

The classes keep the reference's names, argument meaning and error behaviour (dlimg::Exception ->
`dlimgedit_b200.Exception`).  There is NO CPU / PyTorch fallback: importing works without the built
library (so CPU-only tooling can import the package), but every call fails loudly if
libdlimgedit.so is missing or no Blackwell GPU is present.
"""
from __future__ import annotations

import ctypes
import weakref
import enum
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdlimgedit.so")
WEIGHT_FILE_NAME = "mobile_sam_b200.bin"

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_i32p = ctypes.POINTER(ctypes.c_int)


class Exception(RuntimeError):  # noqa: A001 - mirrors dlimg::Exception (dlimgedit.hpp:184)
    pass


class Channels(enum.IntEnum):  # dlimgedit.hpp:29
    mask = 1
    rgb = 3
    rgba = 4
    bgra = 5
    argb = 6


def count(c: Channels) -> int:  # dlimgedit.impl.hpp:15
    return 4 if int(c) > 4 else int(c)


class Backend(enum.IntEnum):  # dlimgedit.hpp:88
    cpu = 0
    gpu = 1


@dataclass
class Extent:
    width: int = 0
    height: int = 0


@dataclass
class Point:  # dlimgedit.hpp:119
    x: int = 0
    y: int = 0


@dataclass
class Region:  # dlimgedit.hpp:125-134
    top_left: Point
    bottom_right: Point

    @staticmethod
    def from_origin(origin: Point, extent: Extent) -> "Region":
        return Region(origin, Point(origin.x + extent.width, origin.y + extent.height))


# ---- C structures (include/dlimg_b200.h) -------------------------------------------------------
class _ImageView(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("channels", ctypes.c_int),
                ("stride", ctypes.c_int), ("pixels", ctypes.c_void_p)]


class _Options(ctypes.Structure):
    _fields_ = [("backend", ctypes.c_int), ("model_directory", ctypes.c_char_p)]


class _Prompt(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("x0", ctypes.c_int), ("y0", ctypes.c_int), ("x1", ctypes.c_int),
                ("y1", ctypes.c_int)]


class _ProfileEntry(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 32), ("launches", ctypes.c_uint64), ("ms", ctypes.c_double),
                ("flops", ctypes.c_double), ("bytes", ctypes.c_double)]


class _Stats(ctypes.Structure):
    _fields_ = [("kernel_launches", ctypes.c_uint64), ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64)]


_R = ctypes.c_int  # dlimg_Result
_H = ctypes.c_void_p  # opaque handles
_F = ctypes.CFUNCTYPE


class _Api(ctypes.Structure):  # 13 slots, order of dlimgedit.h:44-68
    _fields_ = [
        ("is_backend_supported", _F(ctypes.c_int, ctypes.c_int)),
        ("create_environment", _F(_R, ctypes.POINTER(_H), ctypes.POINTER(_Options))),
        ("destroy_environment", _F(None, _H)),
        ("process_image_for_segmentation", _F(_R, ctypes.POINTER(_H), ctypes.POINTER(_ImageView), _H)),
        ("get_segmentation_mask", _F(_R, _H, c_i32p, c_i32p, ctypes.POINTER(ctypes.c_void_p), c_f32p)),
        ("get_segmentation_extent", _F(None, _H, c_i32p)),
        ("destroy_segmentation", _F(None, _H)),
        ("segment_objects", _F(_R, ctypes.POINTER(_ImageView), ctypes.c_void_p, _H)),
        ("load_image", _F(_R, ctypes.c_char_p, c_i32p, c_i32p, ctypes.POINTER(ctypes.c_void_p))),
        ("save_image", _F(_R, ctypes.POINTER(_ImageView), ctypes.c_char_p)),
        ("create_image", _F(ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int)),
        ("destroy_image", _F(None, ctypes.c_void_p)),
        ("last_error", _F(ctypes.c_char_p)),
    ]


class _Ext(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("abi_version", ctypes.c_uint32),
        ("set_stream", _F(_R, _H, ctypes.c_void_p)),
        ("synchronize", _F(_R, _H)),
        ("get_stats", _F(_R, _H, ctypes.POINTER(_Stats))),
        ("process_batch", _F(_R, _H, ctypes.POINTER(_ImageView), ctypes.c_int, ctypes.c_int, ctypes.POINTER(_H))),
        ("compute_masks_batch", _F(_R, _H, ctypes.POINTER(_H), ctypes.POINTER(_Prompt), ctypes.c_int, ctypes.c_int,
                                   ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_int)),
        ("get_embedding", _F(_R, _H, c_f32p)),
        ("get_low_res_logits", _F(_R, _H, ctypes.POINTER(_Prompt), c_f32p, c_f32p)),
        ("resize_longest_side", _F(_R, _H, ctypes.POINTER(_ImageView), ctypes.c_int, ctypes.c_void_p, c_i32p)),
        ("image_tensor", _F(_R, _H, ctypes.POINTER(_ImageView), ctypes.c_void_p)),
        ("mask_postprocess", _F(_R, _H, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p)),
        ("threshold_mask", _F(_R, _H, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                              ctypes.c_void_p)),
        ("profile_enable", _F(_R, _H, ctypes.c_int)),
        ("profile_read", _F(_R, _H, ctypes.POINTER(_ProfileEntry), ctypes.c_int, c_i32p)),
        ("get_embedding_async", _F(_R, _H, c_f32p)),
        ("get_embedding_f16_async", _F(_R, _H, ctypes.c_void_p)),
    ]


class _Debug(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("act_is_bf16", ctypes.c_uint32),
        ("gemm", _F(_R, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                    ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                    ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p)),
        ("encode_tap", _F(_R, _H, ctypes.POINTER(_ImageView), ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p,
                          ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t))),
        ("resize_plan", _F(ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_i32p, c_f32p)),
        ("srgb_tables", _F(None, c_f32p, c_f32p)),
        ("window_attention", _F(_R, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p)),
        ("window_attention_simt", _F(_R, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_void_p)),
        ("layernorm_stats", _F(_R, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                               ctypes.c_void_p)),
        ("mlp_fused", _F(_R, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p)),
        ("conv3x3", _F(_R, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p)),
        ("local_conv", _F(_R, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int)),
        ("image_pool_stats", _F(None, ctypes.POINTER(ctypes.c_uint64))),
    ]


EXPORTED_SYMBOLS = ("dlimg_init", "dlimg_b200_ext_init", "dlimg_b200_debug_init")

_lib = None
_api: Optional[_Api] = None
_ext: Optional[_Ext] = None
_dbg: Optional[_Debug] = None


def load_library():
    """dlopen libdlimgedit.so and resolve the tables (the dynamic-loading route of dlimgedit.hpp:178-181)."""
    global _lib, _api, _ext, _dbg
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Exception(f"{LIB_PATH} is missing: build it with `python -m dlimgedit_b200._build` "
                            "(there is no CPU or PyTorch fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        lib.dlimg_init.restype = ctypes.POINTER(_Api)
        lib.dlimg_b200_ext_init.restype = ctypes.POINTER(_Ext)
        lib.dlimg_b200_debug_init.restype = ctypes.POINTER(_Debug)
        _api = lib.dlimg_init().contents
        _ext = lib.dlimg_b200_ext_init().contents
        _dbg = lib.dlimg_b200_debug_init().contents
        assert _ext.struct_size == ctypes.sizeof(_Ext), "dlimg_b200_Ext layout mismatch"
        assert _dbg.struct_size == ctypes.sizeof(_Debug), "dlimg_b200_Debug layout mismatch"
        _lib = lib
    return _lib


def api() -> _Api:
    load_library()
    return _api


def ext() -> _Ext:
    load_library()
    return _ext


def debug() -> _Debug:
    load_library()
    return _dbg


def _check(result: int):
    if result != 0:
        raise Exception(api().last_error().decode(errors="replace"))


# ---- façade ------------------------------------------------------------------------------------
class ImageView:
    """Read-only view of pixels (dlimgedit.hpp:36-45).  `pixels` is a numpy uint8 array (host) or an
    integer device address when `device=True`; `stride` is in bytes (default width * count(channels))."""

    def __init__(self, pixels, extent: Extent = None, channels: Channels = Channels.rgba, stride: int = 0,
                 device: bool = False):
        self.device = device
        self.channels = Channels(channels)
        if device:
            assert extent is not None
            self._addr = int(pixels)
            self._keep = None
        else:
            arr = np.asarray(pixels)
            assert arr.dtype == np.uint8
            if extent is None:
                extent = Extent(arr.shape[1], arr.shape[0])
            if not stride and arr.ndim >= 2:
                arr = np.ascontiguousarray(arr)
            self._keep = arr
            self._addr = arr.ctypes.data
        self.extent = extent
        self.stride = stride or extent.width * count(self.channels)

    def to_c(self) -> _ImageView:
        return _ImageView(self.extent.width, self.extent.height, int(self.channels), self.stride, self._addr)


class Image:
    """Image that owns packed pixels handed out by the library (dlimgedit.hpp:48-82; dlimgedit.impl.hpp:44-66: the
    constructor takes them from `create_image`, `load` from `load_image`, the destructor gives them back through
    `destroy_image`).  While a GPU environment is alive these buffers are page-locked (csrc/image_pool.hpp), so passing
    an Image to `Segmentation.process` / receiving one from `compute_mask` moves over PCIe without pageable staging.
    `pixels` is a numpy view (H, W, C) -- or (H, W) for masks -- that keeps the Image alive."""

    def __init__(self, extent: Extent, channels: Channels = Channels.rgba, _pixels: int = 0):
        self._extent = extent
        self._channels = Channels(channels)
        addr = _pixels or api().create_image(extent.width, extent.height, count(self._channels))
        if not addr:
            raise Exception("create_image failed")
        self._addr = addr
        n = count(self._channels)
        shape = (extent.height, extent.width) if n == 1 else (extent.height, extent.width, n)
        buf = (ctypes.c_uint8 * (extent.width * extent.height * n)).from_address(addr)
        # the ctypes buffer owns the pixels: they go back to the library when the Image AND every numpy view of it are gone
        # (no reference cycle, so that happens at once and a cached page-locked block is re-used by the next call)
        buf._release = weakref.finalize(buf, api().destroy_image, addr)
        self._pixels = np.frombuffer(buf, np.uint8).reshape(shape)

    def extent(self) -> Extent:
        return self._extent

    def channels(self) -> Channels:
        return self._channels

    @property
    def pixels(self) -> np.ndarray:
        return self._pixels

    def size(self) -> int:
        return self._extent.width * self._extent.height * count(self._channels)

    def view(self) -> "ImageView":
        return ImageView(self._pixels, self._extent, self._channels)

    @staticmethod
    def load(filepath) -> "Image":
        e = (ctypes.c_int * 2)()
        ch = ctypes.c_int()
        px = ctypes.c_void_p()
        _check(api().load_image(os.fspath(filepath).encode(), e, ctypes.byref(ch), ctypes.byref(px)))
        if ch.value not in (1, 3, 4):  # grey + alpha files: the reference casts the count to a Channels value that does not exist
            api().destroy_image(px)
            raise Exception(f"Failed to load image {os.fspath(filepath)}: {ch.value} channels have no Channels value")
        return Image(Extent(e[0], e[1]), Channels(ch.value), _pixels=px.value)

    @staticmethod
    def save(img, filepath) -> None:
        view = img.view() if isinstance(img, Image) else img
        c = view.to_c()
        _check(api().save_image(ctypes.byref(c), os.fspath(filepath).encode()))


class Options:  # dlimgedit.hpp:91-96
    def __init__(self, backend: Backend = Backend.cpu, model_directory: str = "models"):
        self.backend = backend
        self.model_directory = model_directory


class Environment:
    """dlimgedit.hpp:102-113.  Models are loaded on first use; must outlive its Segmentations."""

    @staticmethod
    def is_supported(backend: Backend) -> bool:
        return api().is_backend_supported(int(backend)) != 0

    def __init__(self, options: Options = None):
        options = options or Options()
        self._h = _H()
        self._dir = options.model_directory.encode()
        opts = _Options(int(options.backend), self._dir)
        _check(api().create_environment(ctypes.byref(self._h), ctypes.byref(opts)))
        self._segs = weakref.WeakSet()  # the environment must outlive its Segmentations (dlimgedit.hpp:98-100)

    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            for s in list(getattr(self, "_segs", ())):  # garbage-collection order must not free the environment first
                s.close()
            api().destroy_environment(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except BaseException:
            pass

    # -- additive extension -----------------------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        _check(ext().set_stream(self._h, ctypes.c_void_p(cuda_stream)))

    def synchronize(self):
        _check(ext().synchronize(self._h))

    def stats(self) -> dict:
        s = _Stats()
        _check(ext().get_stats(self._h, ctypes.byref(s)))
        return {"kernel_launches": s.kernel_launches, "h2d_bytes": s.h2d_bytes, "d2h_bytes": s.d2h_bytes}

    def profile_enable(self, on: bool):
        _check(ext().profile_enable(self._h, int(on)))

    def profile_read(self) -> dict:
        """{kernel category: {launches, ms, flops, bytes}} since the last read (CUDA-event timed per launch)."""
        arr = (_ProfileEntry * 32)()
        n = ctypes.c_int(0)
        _check(ext().profile_read(self._h, arr, 32, ctypes.byref(n)))
        return {arr[i].name.decode(): {"launches": arr[i].launches, "ms": arr[i].ms, "flops": arr[i].flops,
                                       "bytes": arr[i].bytes} for i in range(n.value)}

    def process_batch(self, views: Sequence[ImageView]) -> List["Segmentation"]:
        n = len(views)
        on_device = views[0].device
        assert all(v.device == on_device for v in views)
        arr = (_ImageView * n)(*[v.to_c() for v in views])
        out = (_H * n)()
        r = ext().process_batch(self._h, arr, n, int(on_device), out)
        segs = [Segmentation._wrap(out[i], self) for i in range(n) if out[i]]
        _check(r)
        return segs

    def compute_masks_batch(self, segs: Sequence["Segmentation"], prompts: Sequence, multi: bool = False,
                            masks_out: Sequence[int] = None, ious_out: int = 0, host_out: Sequence[np.ndarray] = None,
                            host_ious: np.ndarray = None, host_async: bool = False):
        """Host mode (masks_out None): returns (list of uint8 arrays (n, H, W), float32 array (count, n)); pass
        `host_out` / `host_ious` to have the results written into your own (ideally page-locked) arrays instead of
        freshly allocated pageable ones; with `host_async` the call returns once the work is queued and the arrays are
        complete after `Environment.synchronize()` (downloads then overlap the decoder of the following calls).
        Device mode: masks_out = device addresses (one per prompt, n*W*H bytes), ious_out = device address."""
        cnt = len(prompts)
        n = 3 if multi else 1
        harr = (_H * cnt)(*[s._h for s in segs])
        parr = (_Prompt * cnt)(*[_to_prompt(p) for p in prompts])
        if masks_out is None:
            outs = host_out if host_out is not None else [np.empty((n, s.extent().height, s.extent().width), np.uint8) for s in segs]
            assert len(outs) == cnt and all(o.dtype == np.uint8 and o.flags["C_CONTIGUOUS"] for o in outs)
            ptrs = (ctypes.c_void_p * cnt)(*[o.ctypes.data for o in outs])
            ious = host_ious if host_ious is not None else np.zeros((cnt, n), np.float32)
            assert ious.dtype == np.float32 and ious.size == cnt * n
            _check(ext().compute_masks_batch(self._h, harr, parr, cnt, int(multi), ptrs,
                                             ctypes.c_void_p(ious.ctypes.data), 2 if host_async else 0))
            return outs, ious
        ptrs = (ctypes.c_void_p * cnt)(*[int(a) for a in masks_out])
        _check(ext().compute_masks_batch(self._h, harr, parr, cnt, int(multi), ptrs, ctypes.c_void_p(ious_out), 1))
        return None


def _to_prompt(p) -> _Prompt:
    if isinstance(p, Point):
        return _Prompt(0, p.x, p.y, 0, 0)
    if isinstance(p, Region):
        return _Prompt(1, p.top_left.x, p.top_left.y, p.bottom_right.x, p.bottom_right.y)
    raise TypeError("prompt must be Point or Region")


class Segmentation:
    """dlimgedit.hpp:138-168: an image embedding that can be queried for masks."""

    def __init__(self):
        self._h = None
        self._env = None

    @staticmethod
    def _wrap(h, env) -> "Segmentation":
        s = Segmentation()
        s._h = _H(h)
        s._env = env
        env._segs.add(s)
        return s

    @staticmethod
    def process(img: ImageView, env: Environment) -> "Segmentation":
        assert not img.device, "Segmentation.process takes host pixels (use Environment.process_batch for device data)"
        h = _H()
        view = img.to_c()
        r = api().process_image_for_segmentation(ctypes.byref(h), ctypes.byref(view), env.handle())
        seg = Segmentation._wrap(h.value, env) if h else None  # the handle is owned even when the call failed
        _check(r)
        return seg

    def extent(self) -> Extent:
        e = (ctypes.c_int * 2)()
        api().get_segmentation_extent(self._h, e)
        return Extent(e[0], e[1])

    def _mask_call(self, point, region, n_masks):
        # like dlimgedit.impl.hpp:146-168: the results are Images of the library (page-locked while the environment lives)
        e = self.extent()
        masks = [Image(e, Channels.mask).pixels for _ in range(n_masks)]
        ptrs = (ctypes.c_void_p * 3)(*[m.ctypes.data for m in masks] + [None] * (3 - n_masks))
        ious = (ctypes.c_float * 3)(0.0, 0.0, 0.0)
        pt = (ctypes.c_int * 2)(point.x, point.y) if point is not None else None
        rg = (ctypes.c_int * 4)(region.top_left.x, region.top_left.y, region.bottom_right.x,
                                region.bottom_right.y) if region is not None else None
        _check(api().get_segmentation_mask(self._h, pt, rg, ptrs, ious))
        return masks, [ious[i] for i in range(3)]

    def compute_mask(self, prompt) -> np.ndarray:
        """Point -> best mask; Region -> mask of the largest object in the box.  uint8 (H, W), 0 / 255."""
        if isinstance(prompt, Point):
            return self._mask_call(prompt, None, 1)[0][0]
        if isinstance(prompt, Region):
            return self._mask_call(None, prompt, 1)[0][0]
        raise TypeError("prompt must be Point or Region")

    def compute_masks(self, point: Point) -> List[Tuple[np.ndarray, float]]:
        masks, ious = self._mask_call(point, None, 3)
        return list(zip(masks, ious))

    # -- additive extension -----------------------------------------------------------------------
    def embedding(self, out: np.ndarray = None) -> np.ndarray:
        """(1, 256, 64, 64) float32, NCHW like the reference's `image_embeddings`; `out` may be a pinned buffer."""
        if out is None:
            out = np.empty((1, 256, 64, 64), np.float32)
        assert out.dtype == np.float32 and out.size == 256 * 64 * 64 and out.flags["C_CONTIGUOUS"]
        _check(ext().get_embedding(self._h, out.ctypes.data_as(c_f32p)))
        return out

    def embedding_async(self, out: np.ndarray) -> None:
        """Queues the embedding read into `out` (page-locked, 256*64*64 float32); complete after Environment.synchronize()."""
        assert out.dtype == np.float32 and out.size == 256 * 64 * 64 and out.flags["C_CONTIGUOUS"]
        _check(ext().get_embedding_async(self._h, out.ctypes.data_as(c_f32p)))

    def embedding_f16_async(self, out: np.ndarray) -> None:
        """The same as half precision: `out` is page-locked, 256*64*64 float16; complete after Environment.synchronize()."""
        assert out.dtype == np.float16 and out.size == 256 * 64 * 64 and out.flags["C_CONTIGUOUS"]
        _check(ext().get_embedding_f16_async(self._h, out.ctypes.data))

    def low_res_logits(self, prompt) -> Tuple[np.ndarray, np.ndarray]:
        logits = np.empty((4, 256, 256), np.float32)
        iou = np.empty((4,), np.float32)
        p = _to_prompt(prompt)
        _check(ext().get_low_res_logits(self._h, ctypes.byref(p), logits.ctypes.data_as(c_f32p), iou.ctypes.data_as(c_f32p)))
        return logits, iou

    def close(self):
        if getattr(self, "_h", None):
            api().destroy_segmentation(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except BaseException:
            pass
