"""Reference-facing behaviour of the drop-in boundary on the GPU (error paths, ownership, threading)."""
import ctypes
import threading

import numpy as np
import pytest

import dlimgedit_b200 as dl
from conftest import synthetic_image

pytestmark = pytest.mark.gpu


def test_backend_supported():
    assert dl.Environment.is_supported(dl.Backend.gpu) is True
    assert dl.Environment.is_supported(dl.Backend.cpu) is False


def test_missing_weights_fail_loudly(tmp_path):
    (tmp_path / "segmentation").mkdir()
    e = dl.Environment(dl.Options(dl.Backend.gpu, str(tmp_path)))  # models load lazily (dlimgedit.hpp:109)
    with pytest.raises(dl.Exception, match="Could not find model file"):
        dl.Segmentation.process(dl.ImageView(np.zeros((64, 64, 4), np.uint8)), e)
    e.close()


def test_process_errors_leave_destroyable_handle(env):
    api = dl.api()
    h = ctypes.c_void_p()
    bad = dl._ImageView(16, 16, 2, 32, np.zeros(512, np.uint8).ctypes.data)  # channels enum 2 does not exist
    assert api.process_image_for_segmentation(ctypes.byref(h), ctypes.byref(bad), env.handle()) == 1
    assert b"Unsupported channel order" in api.last_error()
    assert h.value  # assigned before the failure, like reference dlimgedit.cpp:55-57
    api.destroy_segmentation(h)
    short = dl._ImageView(16, 16, 4, 8, np.zeros(1024, np.uint8).ctypes.data)  # stride < width*bpp (image.cpp:38)
    assert api.process_image_for_segmentation(ctypes.byref(h), ctypes.byref(short), env.handle()) == 1
    api.destroy_segmentation(h)


def test_compute_mask_requires_a_prompt(env):
    seg = dl.Segmentation.process(dl.ImageView(synthetic_image(300, 200, 3, 5), channels=dl.Channels.rgb), env)
    assert (seg.extent().width, seg.extent().height) == (200, 300)
    out = np.zeros((300, 200), np.uint8)
    ptrs = (ctypes.c_void_p * 3)(out.ctypes.data, None, None)
    ious = (ctypes.c_float * 3)(7.0, 7.0, 7.0)
    assert dl.api().get_segmentation_mask(seg._h, None, None, ptrs, ious) == 1  # ASSERT(point || region)
    pt = (ctypes.c_int * 2)(50, 60)
    assert dl.api().get_segmentation_mask(seg._h, pt, None, ptrs, ious) == 0
    assert list(ious) == [7.0, 7.0, 7.0]  # the single-mask path does not write accuracies (segmentation.cpp:162-165)
    assert set(np.unique(out)) <= {0, 255}


def test_strided_input_equals_packed(env):
    img = synthetic_image(480, 640, 4, 9)
    packed = dl.Segmentation.process(dl.ImageView(img, channels=dl.Channels.rgba), env)
    stride = 640 * 4 + 64
    buf = np.zeros((480, stride), np.uint8)
    buf[:, :2560] = img.reshape(480, -1)
    strided = dl.Segmentation.process(dl.ImageView(buf, dl.Extent(640, 480), dl.Channels.rgba, stride), env)
    assert np.array_equal(packed.embedding(), strided.embedding())
    p = dl.Point(300, 200)
    assert np.array_equal(packed.compute_mask(p), strided.compute_mask(p))


def test_environment_is_thread_safe(env):
    img = synthetic_image(256, 256, 3, 4)
    seg = dl.Segmentation.process(dl.ImageView(img, channels=dl.Channels.rgb), env)
    ref = seg.compute_mask(dl.Point(100, 100))
    results, errors = [], []

    def worker(i):
        try:
            if i % 2 == 0:
                results.append(np.array_equal(seg.compute_mask(dl.Point(100, 100)), ref))
            else:
                s2 = dl.Segmentation.process(dl.ImageView(img, channels=dl.Channels.rgb), env)
                results.append(np.array_equal(s2.compute_mask(dl.Point(100, 100)), ref))
        except BaseException as e:  # noqa
            errors.append(e)

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors and all(results) and len(results) == 8


def test_launch_counter_moves(env):
    before = env.stats()["kernel_launches"]
    dl.Segmentation.process(dl.ImageView(synthetic_image(128, 128, 3, 2), channels=dl.Channels.rgb), env)
    assert env.stats()["kernel_launches"] > before + 50


def test_async_embedding_read_matches_blocking(env):
    """get_embedding_async (download stream) delivers the same bytes as the blocking read once synchronize returns, also
    when several calls with host pixels are in flight (double-buffered upload slots)."""
    import torch
    from conftest import synthetic_image
    imgs = [synthetic_image(1024, 1024, 4, seed=20 + i) for i in range(3)]
    outs = [torch.empty(1, 256, 64, 64).pin_memory() for _ in imgs]
    segs = []
    for im, o in zip(imgs, outs):  # three back-to-back calls, nothing waits in between
        seg = env.process_batch([dl.ImageView(im, channels=dl.Channels.rgba)])[0]
        seg.embedding_async(o.numpy())
        segs.append(seg)
    env.synchronize()
    for seg, o in zip(segs, outs):
        assert np.array_equal(o.numpy(), seg.embedding())
    assert not np.array_equal(outs[0].numpy(), outs[1].numpy())


def test_half_precision_embedding_download(env):
    """get_embedding_f16_async: the fp32 embedding rounded to IEEE half on the device, half the PCIe bytes."""
    import torch
    from conftest import synthetic_image
    seg = env.process_batch([dl.ImageView(synthetic_image(1024, 1024, 4, seed=31), channels=dl.Channels.rgba)])[0]
    h = torch.empty(1, 256, 64, 64, dtype=torch.float16).pin_memory()
    seg.embedding_f16_async(h.numpy())
    env.synchronize()
    assert np.array_equal(h.numpy(), seg.embedding().astype(np.float16))  # round-to-nearest-even on both sides


def test_async_host_masks_match_blocking(env):
    """masks_on_device = 2: host masks whose downloads overlap the following calls; complete after synchronize()."""
    import torch
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (600, 800, 3), dtype=np.uint8)
    seg = dl.Segmentation.process(dl.ImageView(img, channels=dl.Channels.rgb), env)
    prompts = [dl.Point(int(rng.integers(0, 800)), int(rng.integers(0, 600))) for _ in range(12)]
    ref, ref_ious = env.compute_masks_batch([seg] * 12, prompts, multi=False)
    sets = []
    for k in range(3):  # three calls in flight into three sets of page-locked buffers
        hm = [torch.empty(1, 600, 800, dtype=torch.uint8).pin_memory() for _ in range(12)]
        hi = torch.empty(12, 1, dtype=torch.float32).pin_memory()
        env.compute_masks_batch([seg] * 12, prompts, multi=False, host_out=[m.numpy() for m in hm], host_ious=hi.numpy(),
                                host_async=True)
        sets.append((hm, hi))
    env.synchronize()
    for hm, hi in sets:
        for a, b in zip(hm, ref):
            assert np.array_equal(a.numpy(), b)
        assert np.array_equal(hi.numpy(), ref_ious)
    seg.close()
