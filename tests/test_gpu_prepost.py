"""Pre/post-processing kernels against the oracle (oracle/c/prepost_ref.c and torch's bilinear), through the
C ABI.  Integer / byte / index work must be BIT-EXACT (a1, a2-vs-oracle, a3, a8)."""
import ctypes

import numpy as np
import pytest
import torch

import dlimgedit_b200 as dl
from conftest import synthetic_image
from gpu_util import device_view
from oracle import prepost as P

pytestmark = pytest.mark.gpu


def _check(r):
    if r != 0:
        raise dl.Exception(dl.api().last_error().decode())


def _image_tensor(env, img, channels, stride=0):
    t = torch.from_numpy(img).cuda()
    h, w = (img.shape[0], img.shape[1]) if not stride else stride[1:]
    view = device_view(t, channels) if not stride else dl.ImageView(t.data_ptr(), dl.Extent(w, h), channels, stride[0], device=True)
    out = torch.zeros(h, w, 3, device="cuda")
    c = view.to_c()
    _check(dl.ext().image_tensor(env.handle(), ctypes.byref(c), out.data_ptr()))
    env.synchronize()
    return out.cpu().numpy()


def test_image_tensor_kat(env):
    # reference test/test_segmentation.cpp:59-83
    expected = {dl.Channels.rgb: [0, 1, 2, 3, 4, 24], dl.Channels.rgba: [0, 1, 2, 4, 5, 32],
                dl.Channels.bgra: [2, 1, 0, 6, 5, 34], dl.Channels.argb: [1, 2, 3, 5, 6, 33]}
    for ch, exp in expected.items():
        bpp = dl.count(ch)
        img = (np.arange(8 * 6 * bpp) % 256).astype(np.uint8).reshape(6, 8, bpp)
        t = _image_tensor(env, img, ch)
        got = [t[0, 0, 0], t[0, 0, 1], t[0, 0, 2], t[0, 1, 0], t[0, 1, 1], t[1, 0, 0]]
        assert got == [float(v) for v in exp]
        assert np.array_equal(t, P.create_image_tensor(img, int(ch)))


@pytest.mark.parametrize("ch", [dl.Channels.mask, dl.Channels.rgb, dl.Channels.rgba, dl.Channels.bgra, dl.Channels.argb])
def test_image_tensor_matches_oracle_and_honours_stride(env, ch):
    rng = np.random.default_rng(int(ch))
    bpp = dl.count(ch)
    img = rng.integers(0, 256, (37, 53, bpp), dtype=np.uint8)
    ref = P.create_image_tensor(img, int(ch))
    assert np.array_equal(_image_tensor(env, img, ch), ref)
    # padded rows: the engine honours `stride` (the reference ignores it in this function)
    stride = 53 * bpp + 64
    buf = np.zeros((37, stride), np.uint8)
    buf[:, :53 * bpp] = img.reshape(37, -1)
    assert np.array_equal(_image_tensor(env, buf, ch, stride=(stride, 37, 53)), ref)


def test_threshold_mask_kat(env):
    # reference test/test_segmentation.cpp:85-99
    vals = torch.tensor([0.0, 0.0, 0.2, -3.1, 0.0, 5.5, 0.0, 0.7, 0.0, 0.9], device="cuda")
    out = torch.zeros(2, 4, dtype=torch.uint8, device="cuda")
    _check(dl.ext().threshold_mask(env.handle(), vals.data_ptr(), 2, 5, 4, 2, out.data_ptr()))
    env.synchronize()
    assert out.cpu().tolist() == [[0, 0, 255, 0], [255, 0, 255, 0]]
    g = torch.Generator(device="cuda").manual_seed(0)
    big = torch.randn(1, 1, 300, 517, device="cuda", generator=g)
    big[0, 0, 5, 5] = 0.0
    big[0, 0, 6, 6] = -0.0
    out = torch.zeros(290, 500, dtype=torch.uint8, device="cuda")
    _check(dl.ext().threshold_mask(env.handle(), big.data_ptr(), 300, 517, 500, 290, out.data_ptr()))
    env.synchronize()
    assert np.array_equal(out.cpu().numpy(), P.write_mask_image(big.cpu().numpy(), 0, 500, 290))


def _resize(env, img, channels, stride=0, hw=None):
    t = torch.from_numpy(img).cuda()
    h, w = hw if hw else img.shape[:2]
    view = dl.ImageView(t.data_ptr(), dl.Extent(w, h), channels, stride, device=True)
    out = torch.zeros(1024 * 1024 * 4, dtype=torch.uint8, device="cuda")
    ext = (ctypes.c_int * 2)()
    c = view.to_c()
    _check(dl.ext().resize_longest_side(env.handle(), ctypes.byref(c), 1024, out.data_ptr(), ext))
    env.synchronize()
    bpp = dl.count(channels)
    return out[: ext[0] * ext[1] * bpp].cpu().numpy().reshape(ext[1], ext[0], bpp)


def test_resize_reference_kat(env):
    # reference test/test_image.cpp:51-69 shape of check, on the longest-side path: 8x8 -> 4x4 uses max_side 4
    img = np.zeros((8, 8, 4), np.uint8)
    for i in range(64):
        img[i // 8, i % 8] = [255, 4 * (i // 8), 4 * (i % 8), 255]
    t = torch.from_numpy(img).cuda()
    view = device_view(t, dl.Channels.rgba).to_c()
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    ext = (ctypes.c_int * 2)()
    _check(dl.ext().resize_longest_side(env.handle(), ctypes.byref(view), 4, out.data_ptr(), ext))
    env.synchronize()
    r = out.cpu().numpy().reshape(4, 4, 4)
    for i in range(16):
        assert r[i // 4, i % 4].tolist() == [255, 2 + 8 * (i // 4), 2 + 8 * (i % 4), 255]


@pytest.mark.parametrize("h,w,ch", [(1200, 1800, dl.Channels.rgb), (512, 512, dl.Channels.rgba), (333, 777, dl.Channels.bgra),
                                    (700, 300, dl.Channels.mask), (1025, 1024, dl.Channels.rgb), (64, 48, dl.Channels.argb)])
def test_resize_matches_oracle_bit_exact(env, h, w, ch):
    img = synthetic_image(h, w, dl.count(ch), seed=h + w)
    need, ow, oh, _ = P.resize_longest_side(w, h)
    assert need
    got = _resize(env, img, ch)
    assert got.shape == (oh, ow, dl.count(ch))
    assert np.array_equal(got, P.resize_srgb(img, ow, oh))


def test_resize_4k_strided_matches_oracle(env):
    # BASELINE config 4: 3840x2160 RGB / BGRA, packed and with stride = w*bpp + 64
    for ch in (dl.Channels.rgb, dl.Channels.bgra):
        bpp = dl.count(ch)
        img = synthetic_image(2160, 3840, bpp, seed=int(ch))
        ref = P.resize_srgb(img, 1024, 576)
        assert np.array_equal(_resize(env, img, ch), ref)
        stride = 3840 * bpp + 64
        buf = np.zeros((2160, stride), np.uint8)
        buf[:, : 3840 * bpp] = img.reshape(2160, -1)
        assert np.array_equal(_resize(env, buf, ch, stride=stride, hw=(2160, 3840)), ref)


def test_resize_passthrough_when_scale_is_one(env):
    img = synthetic_image(700, 1024, 3, seed=1)
    assert np.array_equal(_resize(env, img, dl.Channels.rgb), img)


@pytest.mark.parametrize("h,w", [(1024, 1024), (1200, 1800), (2160, 3840), (512, 512), (683, 1024), (333, 1001), (1024, 683),
                                 (1024, 1019), (50, 100), (1500, 4097), (1023, 1024), (7, 9000)])
def test_mask_postprocess_matches_torch(env, oracle_sam, h, w):
    """Fused 256->1024 bilinear, crop, ->(h,w) bilinear, >0: against the oracle's two F.interpolate calls."""
    from oracle.mobile_sam_ref import SamOnnxDecoder
    g = torch.Generator().manual_seed(h * 3 + w)
    low = torch.nn.functional.interpolate(torch.randn(2, 1, 24, 24, generator=g), size=(256, 256), mode="bicubic")
    dec = SamOnnxDecoder(oracle_sam, True)
    ref_logits = dec.postprocess(low, torch.tensor([float(h), float(w)]))
    ref = (ref_logits > 0).numpy().astype(np.uint8)[:, 0] * 255
    d_low = low.cuda().contiguous()
    out = torch.zeros(2, h, w, dtype=torch.uint8, device="cuda")
    _check(dl.ext().mask_postprocess(env.handle(), d_low.data_ptr(), 2, w, h, out.data_ptr()))
    env.synchronize()
    got = out.cpu().numpy()
    assert set(np.unique(got)) <= {0, 255}
    mismatch = float((got != ref).mean())
    near_zero = float((ref_logits.abs() < 1e-5).float().mean())
    # only pixels whose logit is within float rounding of 0 may differ
    assert mismatch <= near_zero + 1e-6, (mismatch, near_zero)


def test_resize_matches_committed_golden_vectors(env):
    """tests/golden/prepost_vectors.npz (oracle outputs committed as bytes): the GPU path reproduces them, packed and
    with padded rows, without the oracle library being involved at all."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prepost_vectors.npz"))
    chans = {"shrink_rgb": dl.Channels.rgb, "shrink_rgba": dl.Channels.rgba, "enlarge_rgb": dl.Channels.rgb, "enlarge_mask": dl.Channels.mask}
    for name, ch in chans.items():
        img, ref = g[f"resize_{name}_in"], g[f"resize_{name}_out"]
        h, w, bpp = img.shape
        t = torch.from_numpy(img).cuda()
        out = torch.zeros(ref.size, dtype=torch.uint8, device="cuda")
        ext = (ctypes.c_int * 2)()
        for pad in (0, 12):
            stride = w * bpp + pad
            buf = torch.zeros(h, stride, dtype=torch.uint8, device="cuda")
            buf[:, : w * bpp] = t.view(h, -1)
            v = dl.ImageView(buf.data_ptr(), dl.Extent(w, h), ch, stride, device=True).to_c()
            _check(dl.ext().resize_longest_side(env.handle(), ctypes.byref(v), max(ref.shape[0], ref.shape[1]), out.data_ptr(), ext))
            env.synchronize()
            assert (ext[0], ext[1]) == (ref.shape[1], ref.shape[0]), name
            assert np.array_equal(out.cpu().numpy().reshape(ref.shape), ref), (name, pad)
    for ch in (1, 3, 4, 5, 6):
        assert np.array_equal(_image_tensor(env, g[f"tensor_{ch}_in"], dl.Channels(ch)), g[f"tensor_{ch}_out"])
