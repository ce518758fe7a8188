"""Generates the committed golden vectors of tests/golden/ (run from the repository root in the build container):

    python tests/golden/make_golden.py

  resize_kat.npz       the reference's `Image resize` known-answer test (test/test_image.cpp:51-69): input ramp, the expected
                       4x4 result as the reference test states it, and the oracle's output
  prepost_vectors.npz  oracle/c/prepost_ref.c outputs on small seeded inputs (a1 extents + prompt transform, a2 resize for
                       enlarging / shrinking / strided RGB + RGBA, a3 image tensor for every channel order, a8 threshold)
  decoder_hf.npz       transformers.models.sam.SamMaskDecoder (an independent implementation of the decoder half)
                       with the seed-0 synthetic weights on a seeded embedding + prompt: masks of tokens 1..3 (every second
                       row / column) and IoUs

The files pin the oracle: tests/test_golden.py (CPU) checks the freshly built oracle against them, so a silent change of
the restatement shows up as a diff against committed bytes."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import prepost as P  # noqa: E402


def resize_kat():
    # test/test_image.cpp:51-69: 8x8 RGBA, R = A = 255, G = 8 * row index ramp, B = 8 * column ramp -> 4x4: G = 2 + 8 * row ...
    img = np.zeros((8, 8, 4), np.uint8)
    for y in range(8):
        for x in range(8):
            img[y, x] = [255, 4 * y, 4 * x, 255]
    expected = np.zeros((4, 4, 4), np.uint8)
    for y in range(4):
        for x in range(4):
            expected[y, x] = [255, 2 + 8 * y, 2 + 8 * x, 255]
    got = P.resize_srgb(img, 4, 4)
    np.savez_compressed(os.path.join(HERE, "resize_kat.npz"), input=img, expected=expected, oracle=got)
    return int(np.abs(got.astype(int) - expected.astype(int)).max())


def prepost_vectors():
    rng = np.random.default_rng(2024)
    out = {}
    extents = [(13, 19, 26), (13, 19, 10), (19, 13, 26), (1800, 1200, 1024), (3840, 2160, 1024), (512, 512, 1024), (1024, 683, 1024)]
    out["extent_in"] = np.array(extents, np.int32)
    out["extent_out"] = np.array([P.resize_longest_side(w, h, m)[1:3] for (w, h, m) in extents], np.int32)
    scales = [P.resize_longest_side(w, h, m)[3] for (w, h, m) in extents]
    out["extent_scale"] = np.array(scales, np.float32)
    coords = np.array([0, 1, 2, 7, 10, 486, 722, 1799], np.int32)
    out["coords"] = coords
    out["coords_scaled"] = np.array([[P.transform_coord(int(c), s) for c in coords] for s in scales], np.int32)
    for name, (h, w, bpp, ow, oh, stride_pad) in {"shrink_rgb": (45, 77, 3, 41, 24, 0), "shrink_rgba": (60, 90, 4, 32, 21, 12),
                                                   "enlarge_rgb": (19, 13, 3, 18, 26, 0), "enlarge_mask": (16, 16, 1, 37, 37, 0)}.items():
        img = rng.integers(0, 256, (h, w, bpp), dtype=np.uint8)
        img[: h // 2] = (img[: h // 2] // 8) * 8  # some flat areas
        res = P.resize_srgb(img, ow, oh)  # (the GPU test also feeds the same pixels with padded rows: same result)
        out[f"resize_{name}_in"] = img
        out[f"resize_{name}_out"] = res
    for ch in (1, 3, 4, 5, 6):
        bpp = P.bytes_per_pixel(ch)
        img = rng.integers(0, 256, (9, 11, bpp), dtype=np.uint8)
        out[f"tensor_{ch}_in"] = img
        out[f"tensor_{ch}_out"] = P.create_image_tensor(img, ch)
    logits = rng.standard_normal((1, 4, 12, 20)).astype(np.float32)
    logits[0, 1, 3, 3] = 0.0
    logits[0, 2, 5, 5] = -0.0
    out["mask_logits"] = logits
    for i in range(4):
        out[f"mask_{i}"] = P.write_mask_image(logits, i, 17, 9)  # extent smaller than the tensor: row stride comes from the tensor
    np.savez_compressed(os.path.join(HERE, "prepost_vectors.npz"), **out)


def decoder_hf():
    from oracle.mobile_sam_ref import SamOnnxDecoder, build_synthetic
    from test_oracle_model import _hf_decoder
    sam = build_synthetic(0)
    hf = _hf_decoder(sam)
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(1, 256, 64, 64, generator=g)
    dec = SamOnnxDecoder(sam, return_single_mask=False)
    cases = [(torch.tensor([[[300.0, 410.0], [0.0, 0.0]]]), torch.tensor([[1.0, -1.0]])),
             (torch.tensor([[[102.0, 63.0], [287.0, 188.0]]]), torch.tensor([[2.0, 3.0]]))]
    masks, ious = [], []
    with torch.no_grad():
        for coords, labels in cases:
            sparse = dec.embed_points(coords, labels)
            dense = dec.embed_masks(torch.zeros(1, 1, 256, 256), torch.zeros(1))
            pe = sam.prompt_encoder.get_dense_pe()
            m_hf, iou_hf = hf(emb, pe, sparse[:, None], dense, multimask_output=True)
            masks.append(m_hf[0, 0, :, ::2, ::2].numpy().astype(np.float32))  # every second row / column: 128 x 128 per token
            ious.append(iou_hf[0, 0].numpy().astype(np.float32))
    np.savez_compressed(os.path.join(HERE, "decoder_hf.npz"), seed=np.array([11]), coords=np.stack([c[0].numpy() for c, _ in cases]),
                        labels=np.stack([l[0].numpy() for _, l in cases]), masks=np.stack(masks), ious=np.stack(ious))


if __name__ == "__main__":
    print("resize KAT: max |oracle - reference expectation| =", resize_kat())
    prepost_vectors()
    decoder_hf()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
