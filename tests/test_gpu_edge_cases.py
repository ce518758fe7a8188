"""Edge cases of the hot path at the reference boundary: degenerate extents, every channel order through the fused
PatchEmbed kernel, 4K inputs (BASELINE configs[3]) end to end, prompts on the image border, many prompts per call."""
import numpy as np
import pytest
import torch

import dlimgedit_b200 as dl
from conftest import synthetic_image
from gpu_util import cosine, iou
from oracle import prepost as P
from oracle.mobile_sam_ref import EncoderWithPreprocess, SamOnnxDecoder

pytestmark = pytest.mark.gpu


def _oracle_embedding(oracle_sam, img, channels):
    h, w = img.shape[:2]
    need, ow, oh, scale = P.resize_longest_side(w, h)
    px = P.resize_srgb(img, ow, oh) if need else img
    x = torch.from_numpy(P.create_image_tensor(px, int(channels)))
    with torch.no_grad():
        return EncoderWithPreprocess(oracle_sam.image_encoder)(x), scale


@pytest.mark.parametrize("h,w,ch", [(1, 1, dl.Channels.rgba), (3, 5, dl.Channels.rgb), (1000, 7, dl.Channels.mask),
                                    (16, 24, dl.Channels.argb), (1024, 1, dl.Channels.bgra)])
def test_degenerate_extents(env, oracle_sam, h, w, ch):
    """One-pixel, one-column and very thin images go through resize (Catmull-Rom up to 1024 on the long side), the
    fused PatchEmbed and the decoder; the embedding still matches the oracle and the mask has the caller's extent."""
    c = dl.count(ch) if hasattr(dl, "count") else (1 if ch == dl.Channels.mask else 3 if ch == dl.Channels.rgb else 4)
    rng = np.random.default_rng(h * 31 + w)
    img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
    seg = dl.Segmentation.process(dl.ImageView(img, channels=ch), env)
    assert (seg.extent().width, seg.extent().height) == (w, h)
    ref, _ = _oracle_embedding(oracle_sam, img, ch)
    assert cosine(torch.from_numpy(seg.embedding()), ref) >= 0.999
    mask = seg.compute_mask(dl.Point(w // 2, h // 2))
    assert mask.shape == (h, w) and set(np.unique(mask)) <= {0, 255}


@pytest.mark.parametrize("ch", [dl.Channels.mask, dl.Channels.rgb, dl.Channels.rgba, dl.Channels.bgra, dl.Channels.argb])
def test_channel_orders_and_unaligned_rows_through_patch_embed(env, oracle_sam, ch):
    """1024-wide input (no resize): the PatchEmbed kernel reads the caller's pixels directly -- aligned 32-bit loads for
    4-byte pixels, byte loads otherwise and for rows that are not 4-byte aligned (odd stride)."""
    c = 1 if ch == dl.Channels.mask else 3 if ch == dl.Channels.rgb else 4
    img = synthetic_image(640, 1024, c, seed=int(ch))
    ref, _ = _oracle_embedding(oracle_sam, img, ch)
    seg = dl.Segmentation.process(dl.ImageView(img, channels=ch), env)
    assert cosine(torch.from_numpy(seg.embedding()), ref) >= 0.999
    stride = 1024 * c + 3  # odd stride: rows lose their alignment
    buf = np.zeros((640, stride), np.uint8)
    buf[:, :1024 * c] = img.reshape(640, 1024 * c)
    d = torch.from_numpy(buf).cuda()
    view = dl.ImageView(d.data_ptr(), dl.Extent(1024, 640), ch, stride, device=True)
    seg2 = env.process_batch([view])[0]
    assert np.array_equal(seg2.embedding(), seg.embedding())


def test_4k_bgra_end_to_end(env, oracle_sam):
    """BASELINE configs[3]: 3840x2160 BGRA -> resize to 1024x576 -> encoder -> decoder -> mask at full resolution."""
    img = synthetic_image(2160, 3840, 4, seed=5)
    seg = dl.Segmentation.process(dl.ImageView(img, channels=dl.Channels.bgra), env)
    ref_emb, scale = _oracle_embedding(oracle_sam, img, dl.Channels.bgra)
    assert cosine(torch.from_numpy(seg.embedding()), ref_emb) >= 0.999
    dec = SamOnnxDecoder(oracle_sam, return_single_mask=True)
    for p in [dl.Point(1900, 1000), dl.Point(0, 0), dl.Point(3839, 2159)]:
        c, l = P.prompt_tensors((p.x, p.y), None, scale)
        with torch.no_grad():
            masks, _, _ = dec(ref_emb, torch.from_numpy(c), torch.from_numpy(l), torch.zeros(1, 1, 256, 256), torch.zeros(1),
                              torch.tensor([2160.0, 3840.0]))
        ref = P.write_mask_image(masks.numpy(), 0, 3840, 2160)
        got = seg.compute_mask(p)
        assert got.shape == (2160, 3840)
        assert iou(got, ref) >= 0.99, (p, iou(got, ref))


def test_prompt_sweep_1_to_256(env):
    """BASELINE configs[2]: 1..256 prompts per cached embedding; a prompt's mask does not depend on its batch."""
    img = synthetic_image(768, 1024, 3, seed=9)
    seg = dl.Segmentation.process(dl.ImageView(img, channels=dl.Channels.rgb), env)
    rng = np.random.default_rng(2)
    prompts = []
    for i in range(256):
        if i % 4 == 3:
            x0, y0 = int(rng.integers(0, 1000)), int(rng.integers(0, 740))
            prompts.append(dl.Region(dl.Point(x0, y0), dl.Point(x0 + int(rng.integers(16, 1024 - x0)) - 1, y0 + int(rng.integers(16, 768 - y0)) - 1)))
        else:
            prompts.append(dl.Point(int(rng.integers(0, 1024)), int(rng.integers(0, 768))))
    full, ious = env.compute_masks_batch([seg] * 256, prompts, multi=False)
    assert len(full) == 256 and ious.shape == (256, 1)
    for n in (1, 2, 64):
        part, pi = env.compute_masks_batch([seg] * n, prompts[:n], multi=False)
        for k in range(n):
            assert np.array_equal(part[k], full[k])
        assert np.allclose(pi, ious[:n], atol=1e-6)
    for k in (0, 3, 255):
        assert np.array_equal(full[k][0], seg.compute_mask(prompts[k]))


def test_repeatable_bits_across_runs_and_batch_slots(env):
    """Race check by repetition (compute-sanitizer is closed on this pool): the same images give the same bits on every
    run, in every batch slot and next to different neighbours -- the persistent kernels hand tiles, mbarrier phases and
    shared-memory buffers from one image to the next, so a synchronisation slip would show up here."""
    imgs = [synthetic_image(1024, 1024, 4, seed=40 + i) for i in range(5)]
    dev = [torch.from_numpy(im).cuda() for im in imgs]

    def run(order):
        views = [dl.ImageView(dev[i].data_ptr(), dl.Extent(1024, 1024), dl.Channels.rgba, device=True) for i in order]
        segs = env.process_batch(views)
        env.synchronize()
        return [s.embedding() for s in segs]

    base = run([0, 1, 2, 3, 4])
    for _ in range(3):
        again = run([0, 1, 2, 3, 4])
        for a, b in zip(base, again):
            assert np.array_equal(a, b)
    shuffled = run([4, 2, 0, 3, 1, 0, 0])
    for slot, i in enumerate([4, 2, 0, 3, 1, 0, 0]):
        assert np.array_equal(shuffled[slot], base[i])
    # decoder: the same prompts, repeated and reordered
    seg = env.process_batch([dl.ImageView(dev[0].data_ptr(), dl.Extent(1024, 1024), dl.Channels.rgba, device=True)])[0]
    prompts = [dl.Point(100 + 37 * k, 900 - 29 * k) for k in range(20)]
    m0, i0 = env.compute_masks_batch([seg] * 20, prompts, multi=False)
    m1, i1 = env.compute_masks_batch([seg] * 20, prompts[::-1], multi=False)
    for k in range(20):
        assert np.array_equal(m0[k], m1[19 - k]) and i0[k, 0] == i1[19 - k, 0]
