"""Helpers shared by the -m gpu parity tests: everything goes through the C ABI of libdlimgedit.so."""
import ctypes

import numpy as np
import torch

import dlimgedit_b200 as dl


def act_dtype():
    """torch dtype of the engine's 16-bit activations / GEMM operands (fp16 by default, csrc/act.hpp)."""
    return torch.bfloat16 if dl.debug().act_is_bf16 else torch.float16


def gemm(a, b, bias=None, residual=None, row_map=None, act=0, out_f32=False, simt=False, out_rows=None, out=None,
         ln_stats=None, ln_parts=0, stats_out=None):
    """a (M,K), b (N,K) CUDA tensors, both act_dtype() or both fp32 (tf32 path).  Returns the output tensor."""
    M, K = a.shape
    N = b.shape[0]
    tf32 = a.dtype == torch.float32
    rows = out_rows if out_rows is not None else M
    if out is None:
        out = torch.zeros(rows, N, device="cuda", dtype=torch.float32 if out_f32 else act_dtype())
    r = dl.debug().gemm(None, int(tf32), int(simt), a.data_ptr(), b.data_ptr(), M, N, K,
                        bias.data_ptr() if bias is not None else None,
                        residual.data_ptr() if residual is not None else None,
                        row_map.data_ptr() if row_map is not None else None, act, int(out_f32), out.data_ptr(),
                        ln_stats.data_ptr() if ln_stats is not None else None, ln_parts,
                        stats_out.data_ptr() if stats_out is not None else None)
    if r != 0:
        raise dl.Exception(dl.api().last_error().decode())
    torch.cuda.synchronize()
    return out


def device_view(t: torch.Tensor, channels, stride=0):
    """t: CUDA uint8 tensor (h, w, bpp) (or a strided buffer with explicit stride) -> dl.ImageView on the device."""
    h, w = t.shape[0], t.shape[1]
    return dl.ImageView(t.data_ptr(), dl.Extent(w, h), channels, stride, device=True)


def encode_tap(env, imgs, channels, name, numel):
    """imgs: list of CUDA uint8 tensors (h, w, bpp).  Returns the named activation as a flat fp32 CUDA tensor."""
    views = (dl._ImageView * len(imgs))(*[device_view(t, channels).to_c() for t in imgs])
    out = torch.zeros(numel, device="cuda", dtype=torch.float32)
    written = ctypes.c_size_t(0)
    r = dl.debug().encode_tap(env.handle(), views, len(imgs), name.encode(), out.data_ptr(), numel, ctypes.byref(written))
    if r != 0:
        raise dl.Exception(dl.api().last_error().decode())
    assert written.value == numel, (name, written.value, numel)
    return out


def cosine(a, b):
    a = a.flatten().double()
    b = b.flatten().double()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def iou(a, b):
    a = np.asarray(a) > 0
    b = np.asarray(b) > 0
    u = (a | b).sum()
    return 1.0 if u == 0 else float((a & b).sum()) / float(u)
