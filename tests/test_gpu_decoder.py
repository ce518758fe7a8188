"""Prompt encoder + two-way mask decoder + post-processing against the oracle (same synthetic weights).

The decoder is fed the ENGINE's embedding in both arms so that decoder error is isolated from encoder error;
the end-to-end mask IoU test then uses the oracle's own embedding (bar: IoU >= 0.99, BASELINE north_star)."""
import numpy as np
import pytest
import torch

import dlimgedit_b200 as dl
from conftest import synthetic_image
from gpu_util import iou
from oracle import prepost as P
from oracle.mobile_sam_ref import EncoderWithPreprocess, SamOnnxDecoder

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def seg_case(env, oracle_sam):
    img = synthetic_image(1200, 1800, 3, seed=21)  # truck-sized RGB (reference test fixture extent)
    seg = dl.Segmentation.process(dl.ImageView(img, channels=dl.Channels.rgb), env)
    emb = torch.from_numpy(seg.embedding())
    _, rw, rh, scale = P.resize_longest_side(1800, 1200)
    return img, seg, emb, scale


PROMPTS = [dl.Point(486, 722), dl.Point(320, 210), dl.Point(1700, 60), dl.Point(0, 0),
           dl.Region(dl.Point(180, 110), dl.Point(505, 330)), dl.Region(dl.Point(900, 400), dl.Point(1500, 1100))]


def _oracle_prompt(p, scale):
    if isinstance(p, dl.Point):
        c, l = P.prompt_tensors((p.x, p.y), None, scale)
    else:
        c, l = P.prompt_tensors(None, (p.top_left.x, p.top_left.y, p.bottom_right.x, p.bottom_right.y), scale)
    return torch.from_numpy(c), torch.from_numpy(l)


@pytest.mark.parametrize("pi", range(len(PROMPTS)))
def test_low_res_logits_and_iou(env, oracle_sam, seg_case, pi):
    _, seg, emb, scale = seg_case
    p = PROMPTS[pi]
    dec = SamOnnxDecoder(oracle_sam, return_single_mask=False)
    c, l = _oracle_prompt(p, scale)
    with torch.no_grad():
        ref, ref_iou = dec.low_res(emb, c, l)
    got, got_iou = seg.low_res_logits(p)
    got = torch.from_numpy(got)
    scale_ref = float(ref.abs().mean())
    err = float((got - ref[0]).abs().max())
    print(f"prompt {pi}: max|dlogit| {err:.5f} (mean|logit| {scale_ref:.3f}); iou {got_iou} vs {ref_iou[0].tolist()}")
    assert err < 2e-2 * max(1.0, scale_ref)  # tf32 image-side GEMMs, fp32 everywhere else
    assert np.allclose(got_iou, ref_iou[0].numpy(), atol=5e-3)
    agree = float(((got > 0) == (ref[0] > 0)).float().mean())
    assert agree > 0.999


@pytest.mark.parametrize("pi", range(len(PROMPTS)))
def test_masks_match_oracle(env, oracle_sam, seg_case, pi):
    """compute_mask through the reference API vs the oracle's single-mask graph + write_mask_image."""
    _, seg, emb, scale = seg_case
    p = PROMPTS[pi]
    dec = SamOnnxDecoder(oracle_sam, return_single_mask=True)
    c, l = _oracle_prompt(p, scale)
    with torch.no_grad():
        masks, _, _ = dec(emb, c, l, torch.zeros(1, 1, 256, 256), torch.zeros(1), torch.tensor([1200.0, 1800.0]))
    ref = P.write_mask_image(masks.numpy(), 0, 1800, 1200)
    got = seg.compute_mask(p)
    assert got.shape == (1200, 1800) and got.dtype == np.uint8 and set(np.unique(got)) <= {0, 255}
    v = iou(got, ref)
    near0 = float((masks.abs() < 1e-2).float().mean())
    print(f"prompt {pi}: IoU {v:.5f}, coverage {float((ref > 0).mean()):.3f}, |logit|<1e-2 share {near0:.4f}")
    assert v >= 0.99


def test_compute_masks_multi(env, oracle_sam, seg_case):
    _, seg, emb, scale = seg_case
    p = PROMPTS[1]
    dec = SamOnnxDecoder(oracle_sam, return_single_mask=False)
    c, l = _oracle_prompt(p, scale)
    with torch.no_grad():
        masks, ious, _ = dec(emb, c, l, torch.zeros(1, 1, 256, 256), torch.zeros(1), torch.tensor([1200.0, 1800.0]))
    res = seg.compute_masks(p)
    assert len(res) == 3
    for i, (m, acc) in enumerate(res):  # reference uses graph outputs 1..3 (segmentation.cpp:166-173)
        ref = P.write_mask_image(masks.numpy(), i + 1, 1800, 1200)
        assert iou(m, ref) >= 0.99
        assert abs(acc - float(ious[0, i + 1])) < 5e-3
    # the single-mask graph returns the best of tokens 1..3 (identical LFS hashes, Appendix C)
    best = int(np.argmax([acc for _, acc in res]))
    assert np.array_equal(seg.compute_mask(p), res[best][0])


def test_end_to_end_iou_against_full_oracle(env, oracle_sam):
    """Full path on both sides: oracle resize -> tensor -> encoder -> decoder -> mask, vs the engine."""
    img = synthetic_image(512, 512, 4, seed=33)  # cat_and_hat-sized RGBA stand-in (512 -> 1024 Catmull-Rom path)
    need, ow, oh, scale = P.resize_longest_side(512, 512)
    x = torch.from_numpy(P.create_image_tensor(P.resize_srgb(img, ow, oh), int(dl.Channels.rgba)))
    enc = EncoderWithPreprocess(oracle_sam.image_encoder)
    dec = SamOnnxDecoder(oracle_sam, return_single_mask=True)
    seg = dl.Segmentation.process(dl.ImageView(img, channels=dl.Channels.rgba), env)
    with torch.no_grad():
        emb = enc(x)
    worst = 1.0
    for p in [dl.Point(320, 210), dl.Point(220, 355), dl.Region(dl.Point(180, 110), dl.Point(505, 330))]:
        c, l = _oracle_prompt(p, scale)
        with torch.no_grad():
            masks, _, _ = dec(emb, c, l, torch.zeros(1, 1, 256, 256), torch.zeros(1), torch.tensor([512.0, 512.0]))
        ref = P.write_mask_image(masks.numpy(), 0, 512, 512)
        v = iou(seg.compute_mask(p), ref)
        print(f"end-to-end IoU {v:.5f} coverage {float((ref > 0).mean()):.3f}")
        worst = min(worst, v)
    assert worst >= 0.99


def test_batched_prompts_equal_single_calls(env, seg_case):
    _, seg, _, _ = seg_case
    rng = np.random.default_rng(1)
    prompts = [dl.Point(int(rng.integers(0, 1800)), int(rng.integers(0, 1200))) for _ in range(37)]
    prompts += [dl.Region(dl.Point(100, 100), dl.Point(400, 300))] * 3
    masks, ious = env.compute_masks_batch([seg] * len(prompts), prompts, multi=False)
    for i in (0, 5, 36, 38):
        assert np.array_equal(masks[i][0], seg.compute_mask(prompts[i]))
    m3, i3 = env.compute_masks_batch([seg] * 4, prompts[:4], multi=True)
    single = seg.compute_masks(prompts[2])
    for k in range(3):
        assert np.array_equal(m3[2][k], single[k][0]) and abs(i3[2, k] - single[k][1]) < 1e-6


def test_mask_modes_agree(env, seg_case):
    """One decoder, three ways to ask: all four logit planes (get_low_res_logits), the selected plane (compute_mask) and planes
    1..3 (compute_masks) come from the same fused epilogue with a different MaskMode -- post-processing the planes of the first
    by hand must give the masks of the other two, and the selection must follow the exported graph's rule (two prompt points:
    mask 0 is penalised by 500, the best of the rest wins)."""
    _, seg, _, _ = seg_case
    for p in (PROMPTS[0], PROMPTS[1], PROMPTS[2]):  # compute_masks takes points only, like the reference
        logits, iou4 = seg.low_res_logits(p)
        score = iou4 + (2 - 2.5) * np.array([1000.0, 0, 0, 0], np.float32)
        best = int(np.argmax(score))
        assert best in (1, 2, 3)
        d_low = torch.from_numpy(logits).cuda().contiguous()
        out = torch.zeros(4, 1200, 1800, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        assert dl.ext().mask_postprocess(env.handle(), d_low.data_ptr(), 4, 1800, 1200, out.data_ptr()) == 0
        env.synchronize()
        planes = out.cpu().numpy()
        assert np.array_equal(planes[best], seg.compute_mask(p))
        multi = seg.compute_masks(p)
        for k in range(3):
            assert np.array_equal(planes[1 + k], multi[k][0])
            assert abs(multi[k][1] - float(iou4[1 + k])) < 1e-6


def test_device_resident_masks(env, seg_case):
    _, seg, _, _ = seg_case
    prompts = [PROMPTS[0], PROMPTS[4]]
    out = torch.zeros(2, 1200, 1800, dtype=torch.uint8, device="cuda")
    ious = torch.zeros(2, device="cuda")
    env.compute_masks_batch([seg, seg], prompts, multi=False, masks_out=[out[0].data_ptr(), out[1].data_ptr()],
                            ious_out=ious.data_ptr())
    env.synchronize()
    for i, p in enumerate(prompts):
        assert np.array_equal(out[i].cpu().numpy(), seg.compute_mask(p))
    assert float(ious.min()) != 0.0
