"""Parity hardening beyond the benign seed-0 weights (VERDICT r1 "what's weak" 1-2):

  * a STRESS weight set (heavy-tailed BatchNorm / LayerNorm scales, strong residual branches: the encoder trunk reaches
    1e2..1e3) must still meet the north-star bars -- embedding cosine >= 0.999, masks IoU >= 0.99 -- with 16-bit storage;
  * a second seed of the benign set;
  * the reference's one real fixture, test/input/truck.jpg (tests/golden/truck.jpg), through process + the reference's
    own prompts (test/test_segmentation.cpp:139, README.md:29) against the oracle, decoded by the library's load_image;
  * prompts of DIFFERENT images in one decoder pass (BASELINE config 5's shape) equal the per-image calls bit for bit;
  * encoding on one stream and decoding on another (stream switch) is ordered correctly.
All through the C ABI."""
import ctypes
import os

import numpy as np
import pytest
import torch

import dlimgedit_b200 as dl
from conftest import synthetic_image
from gpu_util import cosine, iou
from oracle import prepost as P
from oracle.mobile_sam_ref import EncoderWithPreprocess, SamOnnxDecoder, build_synthetic

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _env_for(tmp_path_factory, seed, stress):
    from dlimgedit_b200 import synthetic_weights
    d = tmp_path_factory.mktemp(f"models_s{seed}_{int(stress)}")
    synthetic_weights.write_model_dir(str(d), seed=seed, stress=stress)
    return dl.Environment(dl.Options(dl.Backend.gpu, str(d)))


def _oracle_masks(sam, emb, prompt, scale, w, h):
    if isinstance(prompt, dl.Point):
        c, l = P.prompt_tensors((prompt.x, prompt.y), None, scale)
    else:
        c, l = P.prompt_tensors(None, (prompt.top_left.x, prompt.top_left.y, prompt.bottom_right.x, prompt.bottom_right.y), scale)
    dec = SamOnnxDecoder(sam, return_single_mask=True)
    with torch.no_grad():
        masks, _, _ = dec(emb, torch.from_numpy(c), torch.from_numpy(l), torch.zeros(1, 1, 256, 256), torch.zeros(1),
                          torch.tensor([float(h), float(w)]))
    return P.write_mask_image(masks.numpy(), 0, w, h), masks


def _oracle_embedding(sam, img, channels):
    h, w = img.shape[:2]
    need, ow, oh, scale = P.resize_longest_side(w, h)
    x = P.resize_srgb(img, ow, oh) if need else img
    t = torch.from_numpy(P.create_image_tensor(x, int(channels)))
    with torch.no_grad():
        return EncoderWithPreprocess(sam.image_encoder)(t), scale


@pytest.mark.parametrize("seed,stress", [(0, True), (1, False), (1, True)])
def test_other_weight_sets_meet_the_bars(tmp_path_factory, seed, stress):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    sam = build_synthetic(seed, stress)
    env = _env_for(tmp_path_factory, seed, stress)
    try:
        img = synthetic_image(768, 1024, 4, seed=40 + seed)
        ref_emb, scale = _oracle_embedding(sam, img, dl.Channels.rgba)
        seg = dl.Segmentation.process(dl.ImageView(img, channels=dl.Channels.rgba), env)
        got = torch.from_numpy(seg.embedding())
        c = cosine(got, ref_emb)
        max_abs = float((got - ref_emb).abs().max())
        print(f"seed {seed} stress {stress}: embedding cosine {c:.6f}, max abs {max_abs:.4f} (oracle std {float(ref_emb.std()):.3f})")
        assert c >= 0.999
        worst = 1.0
        for p in (dl.Point(300, 200), dl.Point(900, 700), dl.Region(dl.Point(100, 100), dl.Point(600, 500))):
            # decoder arm isolated: both sides on the ENGINE's embedding (as tests/test_gpu_decoder.py does) ...
            ref_mask, logits = _oracle_masks(sam, got, p, scale, 1024, 768)
            v = iou(seg.compute_mask(p), ref_mask)
            # ... and end to end against the oracle's own embedding
            ref_mask_e2e, _ = _oracle_masks(sam, ref_emb, p, scale, 1024, 768)
            v2 = iou(seg.compute_mask(p), ref_mask_e2e)
            print(f"   {p}: IoU {v:.5f} (same embedding) / {v2:.5f} (end to end); coverage {float((ref_mask > 0).mean()):.3f}, "
                  f"|logit|<1e-2 share {float((logits.abs() < 1e-2).float().mean()):.4f}")
            worst = min(worst, v, v2)
        assert worst >= 0.99
        seg.close()
    finally:
        env.close()


def _load_truck():
    a = dl.api()
    ext = (ctypes.c_int * 2)()
    ch = ctypes.c_int()
    px = ctypes.c_void_p()
    assert a.load_image(os.path.join(GOLDEN, "truck.jpg").encode(), ext, ctypes.byref(ch), ctypes.byref(px)) == 0, a.last_error()
    img = np.ctypeslib.as_array(ctypes.cast(px, ctypes.POINTER(ctypes.c_uint8)), shape=(ext[1], ext[0], ch.value)).copy()
    a.destroy_image(px)
    return img


def test_truck_fixture_through_the_reference_api(env, oracle_sam):
    """reference test/test_segmentation.cpp:125-150 (`SAM.segmentation[gpu]`, truck.jpg Point{486,722}) -- its golden PNG
    is an LFS stub, so the bar is the oracle on the same pixels: cosine >= 0.999, IoU >= 0.99."""
    img = _load_truck()
    assert img.shape == (1200, 1800, 3)
    ref_emb, scale = _oracle_embedding(oracle_sam, img, dl.Channels.rgb)
    seg = dl.Segmentation.process(dl.ImageView(img, channels=dl.Channels.rgb), env)
    assert (seg.extent().width, seg.extent().height) == (1800, 1200)
    got = torch.from_numpy(seg.embedding())
    c = cosine(got, ref_emb)
    print(f"truck.jpg: embedding cosine {c:.6f}, max abs {float((got - ref_emb).abs().max()):.4f}")
    assert c >= 0.999
    for p in (dl.Point(486, 722), dl.Point(220, 355), dl.Region(dl.Point(180, 110), dl.Point(505, 330))):
        ref_mask, _ = _oracle_masks(oracle_sam, ref_emb, p, scale, 1800, 1200)
        m = seg.compute_mask(p)
        v = iou(m, ref_mask)
        print(f"   {p}: IoU {v:.5f}, coverage {float((m > 0).mean()):.3f}")
        assert v >= 0.99
    ms = seg.compute_masks(dl.Point(486, 722))  # three masks + IoU predictions (dlimgedit.hpp:160-163)
    assert len(ms) == 3 and all(m.shape == (1200, 1800) for m, _ in ms)
    seg.close()


def test_prompts_of_different_images_share_one_pass(env):
    """Config 5's shape: 16 prompts on each of several images in ONE compute_masks_batch call equal the per-image
    calls bit for bit (per-prompt image tables, csrc/engine.cu decode_chunk), also with mixed extents."""
    rng = np.random.default_rng(77)
    imgs = [synthetic_image(1024, 1024, 4, seed=60), synthetic_image(1024, 1024, 4, seed=61), synthetic_image(600, 800, 4, seed=62)]
    segs = [dl.Segmentation.process(dl.ImageView(im, channels=dl.Channels.rgba), env) for im in imgs]
    owners, prompts = [], []
    for k in range(16):  # interleaved: consecutive prompts belong to different images
        for s in segs:
            e = s.extent()
            owners.append(s)
            prompts.append(dl.Point(int(rng.integers(0, e.width)), int(rng.integers(0, e.height))))
    mixed, mixed_iou = env.compute_masks_batch(owners, prompts, multi=False)
    for i, (s, p) in enumerate(zip(owners, prompts)):
        one, one_iou = env.compute_masks_batch([s], [p], multi=False)
        assert np.array_equal(mixed[i], one[0]), i
        assert mixed_iou[i, 0] == one_iou[0, 0]
    multi, multi_iou = env.compute_masks_batch(owners[:6], prompts[:6], multi=True)
    for i in range(6):
        ref = owners[i].compute_masks(prompts[i])
        for k in range(3):
            assert np.array_equal(multi[i][k], ref[k][0])
            assert multi_iou[i, k] == np.float32(ref[k][1])
    for s in segs:
        s.close()


def test_encode_on_one_stream_decode_on_another(env):
    """set_stream orders everything queued on the old stream in front of the new one, and a decoder pass waits for the
    `ready` event of every embedding store it reads (encoding on stream A then decoding on stream B raced)."""
    img = synthetic_image(1024, 1024, 4, seed=70)
    d = torch.from_numpy(img).cuda()
    view = dl.ImageView(d.data_ptr(), dl.Extent(1024, 1024), dl.Channels.rgba, device=True)
    p = dl.Point(500, 400)
    env.set_stream(0)
    ref_seg = env.process_batch([view])[0]
    env.synchronize()
    ref = ref_seg.compute_mask(p)
    a, b = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(5):
        env.set_stream(a.cuda_stream)
        seg = env.process_batch([view] * 4)[3]      # asynchronous: the encoder is still running on stream a ...
        env.set_stream(b.cuda_stream)               # ... when the work stream changes
        m, _ = env.compute_masks_batch([seg], [p], multi=False)
        assert np.array_equal(m[0][0], ref)
        seg.close()                                  # frees the store while stream b may still hold work on it
    env.set_stream(0)
    env.synchronize()
    ref_seg.close()


def test_counters_and_profile_are_per_environment(env, model_dir):
    other = dl.Environment(dl.Options(dl.Backend.gpu, model_dir))
    try:
        before = env.stats()["kernel_launches"]
        dl.Segmentation.process(dl.ImageView(synthetic_image(256, 256, 3, 1), channels=dl.Channels.rgb), other)
        assert other.stats()["kernel_launches"] > 50
        assert env.stats()["kernel_launches"] == before  # the other environment's launches are not counted here
    finally:
        other.close()


def test_many_handles_leave_no_device_or_pinned_memory_behind(env):
    """The reference's usage pattern is a long-lived Environment with one short-lived Segmentation per opened image
    (README.md:20-33).  After a warm-up the device pool and the page-locked image pool must be stable over many
    process -> compute_mask -> destroy cycles of differently sized images."""
    import ctypes
    import torch
    sizes = [(300, 500, 3), (640, 480, 4), (1024, 1024, 4), (720, 1280, 3)]
    imgs = [synthetic_image(h, w, c, seed=20 + i) for i, (h, w, c) in enumerate(sizes)]

    def cycle(i):
        img = imgs[i % len(imgs)]
        ch = dl.Channels.rgb if img.shape[2] == 3 else dl.Channels.rgba
        seg = dl.Segmentation.process(dl.ImageView(img, channels=ch), env)
        m = seg.compute_mask(dl.Point(img.shape[1] // 2, img.shape[0] // 2))
        assert m.shape == img.shape[:2]
        seg.close()

    def pool_in_use():
        import gc
        gc.collect()  # (library buffers that only a reference cycle of an earlier test keeps alive are released first)
        st = (ctypes.c_uint64 * 5)()
        dl.debug().image_pool_stats(st)
        return st[0]

    for i in range(8):
        cycle(i)
    env.synchronize()
    torch.cuda.synchronize()
    free0, pinned0 = torch.cuda.mem_get_info()[0], pool_in_use()
    for i in range(120):
        cycle(i)
    env.synchronize()
    torch.cuda.synchronize()
    free1, pinned1 = torch.cuda.mem_get_info()[0], pool_in_use()
    assert free0 - free1 < (64 << 20), (free0, free1)  # device memory: nothing accumulates
    assert pinned1 == pinned0                          # every mask buffer went back to the pool
