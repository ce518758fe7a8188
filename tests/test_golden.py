"""The oracle against the committed golden vectors (tests/golden/, made by tests/golden/make_golden.py): the C
restatement of the pre/post path and the PyTorch restatement of the decoder are pinned to fixed bytes, so a silent change
of either shows up here.  CPU only."""
import os

import numpy as np
import torch

from oracle import prepost as P

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_reference_resize_kat():
    """reference test/test_image.cpp:51-69, bit for bit."""
    g = np.load(os.path.join(GOLDEN, "resize_kat.npz"))
    got = P.resize_srgb(g["input"], 4, 4)
    assert np.array_equal(got, g["expected"]) and np.array_equal(got, g["oracle"])


def test_prepost_vectors():
    g = np.load(os.path.join(GOLDEN, "prepost_vectors.npz"))
    for (w, h, m), (ow, oh), sc in zip(g["extent_in"], g["extent_out"], g["extent_scale"]):
        _, w2, h2, s2 = P.resize_longest_side(int(w), int(h), int(m))
        assert (w2, h2) == (int(ow), int(oh)) and np.float32(s2) == sc
    for sc, row in zip(g["extent_scale"], g["coords_scaled"]):
        assert [P.transform_coord(int(c), sc) for c in g["coords"]] == row.tolist()
    for name in ("shrink_rgb", "shrink_rgba", "enlarge_rgb", "enlarge_mask"):
        ref = g[f"resize_{name}_out"]
        assert np.array_equal(P.resize_srgb(g[f"resize_{name}_in"], ref.shape[1], ref.shape[0]), ref), name
    for ch in (1, 3, 4, 5, 6):
        assert np.array_equal(P.create_image_tensor(g[f"tensor_{ch}_in"], ch), g[f"tensor_{ch}_out"])
    for i in range(4):
        assert np.array_equal(P.write_mask_image(g["mask_logits"], i, 17, 9), g[f"mask_{i}"])
    assert g["mask_1"][3, 3] == 0 and g["mask_2"][5, 5] == 0  # +0.0 and -0.0 are not > 0


def test_decoder_oracle_matches_transformers_golden(oracle_sam):
    """The decoder half of the oracle against stored outputs of transformers' SamMaskDecoder (independent implementation)."""
    from oracle.mobile_sam_ref import SamOnnxDecoder
    g = np.load(os.path.join(GOLDEN, "decoder_hf.npz"))
    gen = torch.Generator().manual_seed(int(g["seed"][0]))
    emb = torch.randn(1, 256, 64, 64, generator=gen)
    dec = SamOnnxDecoder(oracle_sam, return_single_mask=False)
    for coords, labels, masks, ious in zip(g["coords"], g["labels"], g["masks"], g["ious"]):
        with torch.no_grad():
            low, iou = dec.low_res(emb, torch.from_numpy(coords)[None], torch.from_numpy(labels)[None])
        assert torch.allclose(low[0, 1:, ::2, ::2], torch.from_numpy(masks), atol=3e-4, rtol=1e-4)
        assert torch.allclose(iou[0, 1:], torch.from_numpy(ious), atol=1e-5, rtol=1e-4)
