"""Deterministic synthetic MobileSAM weights (numpy only).

No MobileSAM checkpoint exists offline, so tests, smoke() and bench.py run the engine
on a seeded random-init model of exactly the checkpoint's architecture and tensor names (SURVEY Appendix
A.7).  The scales are chosen so activations stay O(1) through the 40-layer encoder and the mask logits are
not degenerate; every tensor of the real checkpoint (BatchNorm running statistics, LayerNorm affines,
attention biases, ...) gets a non-trivial value so a parity bug in any of them is visible.

    python -m dlimgedit_b200.synthetic_weights <model_dir> [seed]   # writes <model_dir>/segmentation/mobile_sam_b200.bin
"""
from __future__ import annotations

import math
import os
import sys
from typing import Dict

import numpy as np

from . import weights_io

DIMS = (64, 128, 160, 320)
DEPTHS = (2, 2, 6, 2)
HEADS = (2, 4, 5, 10)
WINDOWS = (7, 7, 14, 7)


class _Gen:
    def __init__(self, seed: int, stress: bool = False):
        self.rng = np.random.default_rng(seed)
        self.t: Dict[str, np.ndarray] = {}
        self.stress = stress

    def _affine(self, shape, lo, hi, cap):
        """Norm scale: benign = uniform(lo, hi); stress = the same times a heavy log-normal tail, capped at `cap`."""
        w = self.rng.uniform(lo, hi, shape)
        if self.stress:
            w = np.minimum(w * np.exp(self.rng.standard_normal(shape) * 0.5), cap)
        return w.astype(np.float32)

    def normal(self, name, shape, std):
        self.t[name] = (self.rng.standard_normal(shape) * std).astype(np.float32)

    def uniform(self, name, shape, lo, hi):
        self.t[name] = self.rng.uniform(lo, hi, shape).astype(np.float32)

    def conv_bn(self, p, cout, cin_per_group, ks, gamma_scale=1.0):
        self.normal(p + ".c.weight", (cout, cin_per_group, ks, ks), 1.0 / math.sqrt(cin_per_group * ks * ks))
        self.t[p + ".bn.weight"] = self._affine((cout,), 0.7 * gamma_scale, 1.3 * gamma_scale, 8.0 * gamma_scale)
        self.normal(p + ".bn.bias", (cout,), 0.5 if self.stress else 0.1)
        self.normal(p + ".bn.running_mean", (cout,), 0.1)
        self.uniform(p + ".bn.running_var", (cout,), 0.6, 1.4)

    def linear(self, p, n, k, scale=1.0):
        self.normal(p + ".weight", (n, k), scale / math.sqrt(k))
        self.normal(p + ".bias", (n,), 0.05)

    def norm(self, p, c):
        self.t[p + ".weight"] = self._affine((c,), 0.7, 1.3, 4.0)
        self.normal(p + ".bias", (c,), 0.5 if self.stress else 0.1)

    def attn(self, p, dim, internal):
        for n in ("q_proj", "k_proj", "v_proj"):
            self.linear(f"{p}.{n}", internal, dim)
        self.linear(p + ".out_proj", dim, internal)


def make_state_dict(seed: int = 0, stress: bool = False) -> Dict[str, np.ndarray]:
    """stress = False: every tensor O(1), BN gamma in [0.7, 1.3] (benign).  stress = True: heavy-tailed BN / LN scales
    (up to 8 / 4), larger norm biases and residual branches 2x stronger, so the encoder trunk grows to 1e2..1e3 with a
    large per-token mean -- the regime in which 16-bit activation storage and the E[x^2] - mean^2 LayerNorm statistics
    of the fused epilogues would break first (real TinyViT residual streams are not O(1))."""
    g = _Gen(seed, stress)
    rs = 2.0 if stress else 1.0  # residual-branch scale
    E = "image_encoder."
    g.conv_bn(E + "patch_embed.seq.0", 32, 3, 3)
    g.conv_bn(E + "patch_embed.seq.2", 64, 32, 3)
    for i in range(2):  # layer 0: MBConv x2 (residual branch damped)
        p = f"{E}layers.0.blocks.{i}"
        g.conv_bn(p + ".conv1", 256, 64, 1)
        g.conv_bn(p + ".conv2", 256, 1, 3)
        g.conv_bn(p + ".conv3", 64, 256, 1, gamma_scale=0.5 * rs)
    for i in range(3):  # PatchMerging after layers 0..2
        p = f"{E}layers.{i}.downsample"
        g.conv_bn(p + ".conv1", DIMS[i + 1], DIMS[i], 1)
        g.conv_bn(p + ".conv2", DIMS[i + 1], 1, 3)
        g.conv_bn(p + ".conv3", DIMS[i + 1], DIMS[i + 1], 1)
    for st in range(1, 4):
        c, ws = DIMS[st], WINDOWS[st]
        for i in range(DEPTHS[st]):
            p = f"{E}layers.{st}.blocks.{i}"
            g.norm(p + ".attn.norm", c)
            g.linear(p + ".attn.qkv", 3 * c, c)
            g.linear(p + ".attn.proj", c, c, scale=0.5 * rs)
            g.normal(p + ".attn.attention_biases", (HEADS[st], ws * ws), 0.5)
            g.conv_bn(p + ".local_conv", c, 1, 3)
            g.norm(p + ".mlp.norm", c)
            g.linear(p + ".mlp.fc1", 4 * c, c)
            g.linear(p + ".mlp.fc2", c, 4 * c, scale=0.5 * rs)
    g.norm(E + "norm_head", 320)         # unused by the forward pass, present in the checkpoint
    g.linear(E + "head", 1000, 320)      # idem
    g.normal(E + "neck.0.weight", (256, 320, 1, 1), 1.0 / math.sqrt(320))
    g.norm(E + "neck.1", 256)
    g.normal(E + "neck.2.weight", (256, 256, 3, 3), 1.0 / math.sqrt(256 * 9))
    g.norm(E + "neck.3", 256)

    P = "prompt_encoder."
    g.normal(P + "pe_layer.positional_encoding_gaussian_matrix", (2, 128), 1.0)
    for i in range(4):
        g.normal(f"{P}point_embeddings.{i}.weight", (1, 256), 0.5)
    g.normal(P + "not_a_point_embed.weight", (1, 256), 0.5)
    g.normal(P + "no_mask_embed.weight", (1, 256), 0.5)
    for idx, (co, ci, ks) in {0: (4, 1, 2), 3: (16, 4, 2), 6: (256, 16, 1)}.items():  # never influences the output
        g.normal(f"{P}mask_downscaling.{idx}.weight", (co, ci, ks, ks), 1.0 / math.sqrt(ci * ks * ks))
        g.normal(f"{P}mask_downscaling.{idx}.bias", (co,), 0.05)
    g.norm(P + "mask_downscaling.1", 4)
    g.norm(P + "mask_downscaling.4", 16)

    D = "mask_decoder."
    g.normal(D + "iou_token.weight", (1, 256), 0.5)
    g.normal(D + "mask_tokens.weight", (4, 256), 0.5)
    for i in range(2):
        p = f"{D}transformer.layers.{i}"
        g.attn(p + ".self_attn", 256, 256)
        g.attn(p + ".cross_attn_token_to_image", 256, 128)
        g.attn(p + ".cross_attn_image_to_token", 256, 128)
        for n in range(1, 5):
            g.norm(f"{p}.norm{n}", 256)
        g.linear(p + ".mlp.lin1", 2048, 256)
        g.linear(p + ".mlp.lin2", 256, 2048)
    g.attn(D + "transformer.final_attn_token_to_image", 256, 128)
    g.norm(D + "transformer.norm_final_attn", 256)
    g.normal(D + "output_upscaling.0.weight", (256, 64, 2, 2), 1.0 / math.sqrt(256))
    g.normal(D + "output_upscaling.0.bias", (64,), 0.05)
    g.norm(D + "output_upscaling.1", 64)
    g.normal(D + "output_upscaling.3.weight", (64, 32, 2, 2), 1.0 / math.sqrt(64))
    g.normal(D + "output_upscaling.3.bias", (32,), 0.05)
    for m in range(4):
        for j, (n, k) in enumerate(((256, 256), (256, 256), (32, 256))):
            g.linear(f"{D}output_hypernetworks_mlps.{m}.layers.{j}", n, k)
    for j, (n, k) in enumerate(((256, 256), (256, 256), (4, 256))):
        g.linear(f"{D}iou_prediction_head.layers.{j}", n, k)
    return g.t


def write_model_dir(model_dir: str, seed: int = 0, stress: bool = False) -> str:
    d = os.path.join(model_dir, "segmentation")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, weights_io.WEIGHT_FILE_NAME)
    weights_io.save(path, make_state_dict(seed, stress))
    return path


if __name__ == "__main__":
    print(write_model_dir(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0))
