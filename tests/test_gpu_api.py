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


def _pool():
    import gc
    gc.collect()  # masks of earlier tests that only a reference cycle keeps alive go back to the pool now, not in mid-test
    st = (ctypes.c_uint64 * 5)()
    dl.debug().image_pool_stats(st)
    return dict(in_use=st[0], cached=st[1], pinned_allocs=st[2], reuses=st[3], plain=st[4])


def test_library_images_are_page_locked_and_recycled(env):
    """create_image / load_image hand out page-locked blocks while an environment lives (csrc/image_pool.hpp), destroy_image
    caches them by size; results through them equal the results through plain numpy buffers."""
    import os
    import torch
    truck = os.path.join(os.path.dirname(__file__), "golden", "truck.jpg")
    before = _pool()
    img = dl.Image.load(truck)
    mid = _pool()
    assert mid["in_use"] - before["in_use"] >= img.size()          # the loader's buffer is a pool block
    assert mid["pinned_allocs"] + mid["reuses"] == before["pinned_allocs"] + before["reuses"] + 1
    assert torch.from_numpy(img.pixels).is_pinned()                  # and CUDA agrees that it is page-locked
    seg_a = dl.Segmentation.process(img.view(), env)
    seg_b = dl.Segmentation.process(dl.ImageView(img.pixels.copy(), channels=dl.Channels.rgb), env)  # pageable copy
    assert np.array_equal(seg_a.embedding(), seg_b.embedding())
    m1 = seg_a.compute_mask(dl.Point(486, 722))                      # an Image of the library, like dlimgedit.impl.hpp:146
    assert torch.from_numpy(m1).is_pinned() and m1.shape == (1200, 1800)
    plain = np.empty((1200, 1800), np.uint8)
    ptrs = (ctypes.c_void_p * 3)(plain.ctypes.data, None, None)
    acc = (ctypes.c_float * 3)()
    assert dl.api().get_segmentation_mask(seg_b._h, (ctypes.c_int * 2)(486, 722), None, ptrs, acc) == 0
    assert np.array_equal(m1, plain)
    # release -> cached -> re-used by the next request of the same size
    keep = m1.copy()
    del m1
    a = _pool()
    m2 = seg_a.compute_mask(dl.Point(486, 722))
    b = _pool()
    assert b["reuses"] == a["reuses"] + 1 and b["pinned_allocs"] == a["pinned_allocs"]
    assert np.array_equal(m2, keep)
    small = dl.Image(dl.Extent(16, 16), dl.Channels.rgba)             # below the pinning threshold: plain memory
    assert _pool()["plain"] == b["plain"] + 1
    del small, m2, img
    seg_a.close()
    seg_b.close()
    assert _pool()["in_use"] == before["in_use"]


def test_images_outlive_their_environment(model_dir):
    """Blocks still in use when the last environment goes are released later without one; the cache is dropped with it."""
    base = _pool()
    e = dl.Environment(dl.Options(dl.Backend.gpu, model_dir))
    img = dl.Image(dl.Extent(512, 512), dl.Channels.rgba)
    tmp = dl.Image(dl.Extent(512, 256), dl.Channels.rgba)
    del tmp
    img.pixels[:] = 3
    e.close()
    assert int(img.pixels.sum()) == 3 * 512 * 512 * 4
    del img
    after = _pool()
    assert after["in_use"] == base["in_use"]


def test_neighbouring_page_locked_buffers_are_not_copied_as_one(env):
    """The engine merges the copies of host buffers that follow each other exactly; two SEPARATE page-locked regions
    that happen to be neighbours cannot be spanned by one cudaMemcpyAsync (cudaErrorInvalidValue), so the merged copy
    falls back to one copy per buffer.  Seen with three 1 MiB masks from three neighbouring cudaHostAlloc blocks."""
    import torch
    rt = torch.cuda.cudart()
    n = 1024 * 1024
    raw = np.empty(4 * n + 4096, np.uint8)
    off = (-raw.ctypes.data) % 4096
    base = raw[off:off + 4 * n]
    parts = [base[k * n:(k + 1) * n] for k in range(4)]
    registered = []
    try:
        for p in parts:  # four separate registrations, back to back in the address space
            assert int(rt.cudaHostRegister(p.ctypes.data, n, 0)) == 0
            registered.append(p.ctypes.data)
        rng = np.random.default_rng(5)
        img = synthetic_image(1024, 1024, 4, seed=31)
        # input side: two images in neighbouring registrations go through process_batch as one packed run
        parts[0][:] = 0
        two = [parts[2].reshape(512, 512, 4), parts[3].reshape(512, 512, 4)]
        two[0][:] = img[:512, :512]
        two[1][:] = img[512:, 512:]
        segs2 = env.process_batch([dl.ImageView(t, channels=dl.Channels.rgba) for t in two])
        env.synchronize()
        for t, s2 in zip(two, segs2):
            alone = dl.Segmentation.process(dl.ImageView(t.copy(), channels=dl.Channels.rgba), env)
            assert np.array_equal(s2.embedding(), alone.embedding())
            alone.close()
            s2.close()
        # output side: masks of consecutive prompts into neighbouring registrations
        seg = dl.Segmentation.process(dl.ImageView(img, channels=dl.Channels.rgba), env)
        prompts = [dl.Point(int(rng.integers(0, 1024)), int(rng.integers(0, 1024))) for _ in range(2)]
        ref, _ = env.compute_masks_batch([seg] * 2, prompts, multi=False)
        outs = [parts[0].reshape(1, 1024, 1024), parts[1].reshape(1, 1024, 1024)]
        ious = np.empty((2, 1), np.float32)
        env.compute_masks_batch([seg] * 2, prompts, multi=False, host_out=outs, host_ious=ious)
        for a, b in zip(outs, ref):
            assert np.array_equal(a, b)
        seg.close()
    finally:
        env.synchronize()
        for addr in registered:
            rt.cudaHostUnregister(addr)
