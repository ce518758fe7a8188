"""ORACLE (test infrastructure only) -- fp32 PyTorch-CPU restatement of the MobileSAM graphs that
dlimgedit executes through onnxruntime.

PARITY UNPINNED: the three .onnx graphs, onnxruntime and the golden PNGs are not available offline
(SURVEY.md section 0, 8c).  This file restates the *published* MobileSAM / segment-anything
algorithm that `script/export_models.py:21-43` exports, anchored on the reference call sites:

  * encoder graph input/outputs        reference src/segmentation.cpp:14-16, 35-41
  * decoder graph inputs/outputs       reference src/segmentation.cpp:19-24, 131-174
  * constant mask input / has_mask=0   reference src/segmentation.cpp:43-45

Self-consistency gates (tests/test_oracle_model.py): learnable-parameter count == 10,130,092 with the
published split (prompt encoder 6,220 / mask decoder 4,058,340), and everything but the TinyViT trunk agrees
with `transformers.models.sam` (an independent restatement of the same published algorithm) given copied
weights: mask decoder, prompt assembly + dense positional grid, mask post-processing + threshold, the
resized-extent rule and the normalise + pad preprocessing.  The TinyViT-5M trunk has no independent
implementation in this image (no timm): it is pinned by the parameter split and its shapes only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product path (dlimgedit_b200/csrc) never does.
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

IMG_SIZE = 1024
PIXEL_MEAN = (123.675, 116.28, 103.53)
PIXEL_STD = (58.395, 57.12, 57.375)


# --------------------------------------------------------------------------------------------
# TinyViT-5M image encoder (MobileSAM `vit_t`): SURVEY Appendix A.2 / A.3
# --------------------------------------------------------------------------------------------
class Conv2dBN(nn.Sequential):
    """conv (no bias) + BatchNorm2d; state-dict names `c.weight`, `bn.*` (Appendix A.7)."""

    def __init__(self, cin, cout, ks=1, stride=1, pad=0, groups=1):
        super().__init__()
        self.add_module("c", nn.Conv2d(cin, cout, ks, stride, pad, groups=groups, bias=False))
        self.add_module("bn", nn.BatchNorm2d(cout))


class PatchEmbed(nn.Module):
    def __init__(self, cin, dim):
        super().__init__()
        self.seq = nn.Sequential(Conv2dBN(cin, dim // 2, 3, 2, 1), nn.GELU(), Conv2dBN(dim // 2, dim, 3, 2, 1))

    def forward(self, x):
        return self.seq(x)


class MBConv(nn.Module):
    def __init__(self, cin, cout, expand):
        super().__init__()
        hid = int(cin * expand)
        self.conv1 = Conv2dBN(cin, hid, 1)
        self.conv2 = Conv2dBN(hid, hid, 3, 1, 1, groups=hid)
        self.conv3 = Conv2dBN(hid, cout, 1)
        self.act = nn.GELU()

    def forward(self, x):
        s = x
        x = self.act(self.conv1(x))
        x = self.act(self.conv2(x))
        x = self.conv3(x)
        return self.act(x + s)


class PatchMerging(nn.Module):
    def __init__(self, res, dim, out_dim):
        super().__init__()
        self.res = res
        self.conv1 = Conv2dBN(dim, out_dim, 1)
        stride = 1 if out_dim in (320, 448, 576) else 2  # MobileSAM special case
        self.conv2 = Conv2dBN(out_dim, out_dim, 3, stride, 1, groups=out_dim)
        self.conv3 = Conv2dBN(out_dim, out_dim, 1)
        self.act = nn.GELU()

    def forward(self, x):
        if x.ndim == 3:
            b = x.shape[0]
            x = x.view(b, self.res, self.res, -1).permute(0, 3, 1, 2)
        x = self.act(self.conv1(x))
        x = self.act(self.conv2(x))
        x = self.conv3(x)
        return x.flatten(2).transpose(1, 2)


class ConvLayer(nn.Module):
    def __init__(self, dim, res, depth, out_dim, expand):
        super().__init__()
        self.blocks = nn.ModuleList([MBConv(dim, dim, expand) for _ in range(depth)])
        self.downsample = PatchMerging(res, dim, out_dim)

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return self.downsample(x)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)
        self.act = nn.GELU()

    def forward(self, x):
        return self.fc2(self.act(self.fc1(self.norm(x))))


def attention_bias_idxs(ws: int) -> Tuple[torch.Tensor, int]:
    """First-seen ordering of (|drow|,|dcol|) offsets over itertools.product pairs (Appendix A.3)."""
    pts = list(itertools.product(range(ws), range(ws)))
    offsets: Dict[Tuple[int, int], int] = {}
    idxs = []
    for p1 in pts:
        for p2 in pts:
            o = (abs(p1[0] - p2[0]), abs(p1[1] - p2[1]))
            if o not in offsets:
                offsets[o] = len(offsets)
            idxs.append(offsets[o])
    n = len(pts)
    return torch.tensor(idxs, dtype=torch.long).view(n, n), len(offsets)


class WindowAttention(nn.Module):
    def __init__(self, dim, heads, ws):
        super().__init__()
        self.heads = heads
        self.kd = dim // heads
        self.scale = self.kd ** -0.5
        self.norm = nn.LayerNorm(dim)
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)
        idxs, n_off = attention_bias_idxs(ws)
        self.attention_biases = nn.Parameter(torch.zeros(heads, n_off))
        self.register_buffer("attention_bias_idxs", idxs, persistent=False)

    def forward(self, x):
        b, n, _ = x.shape
        x = self.norm(x)
        qkv = self.qkv(x).view(b, n, self.heads, 3 * self.kd)  # per-head interleaved [q|k|v]
        q, k, v = qkv.split([self.kd, self.kd, self.kd], dim=3)
        q, k, v = (t.permute(0, 2, 1, 3) for t in (q, k, v))
        attn = (q @ k.transpose(-2, -1)) * self.scale + self.attention_biases[:, self.attention_bias_idxs]
        attn = attn.softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(b, n, self.heads * self.kd)
        return self.proj(x)


class TinyViTBlock(nn.Module):
    def __init__(self, dim, res, heads, ws, mlp_ratio=4.0):
        super().__init__()
        self.res, self.ws = res, ws
        self.attn = WindowAttention(dim, heads, ws)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.local_conv = Conv2dBN(dim, dim, 3, 1, 1, groups=dim)

    def forward(self, x):
        h = w = self.res
        b, l, c = x.shape
        ws = self.ws
        res_x = x
        x = x.view(b, h, w, c)
        pad = (ws - h % ws) % ws
        if pad:
            x = F.pad(x, (0, 0, 0, pad, 0, pad))  # zeros BEFORE the in-attention LayerNorm; not masked
        ph = h + pad
        nh = ph // ws
        x = x.view(b, nh, ws, nh, ws, c).transpose(2, 3).reshape(b * nh * nh, ws * ws, c)
        x = self.attn(x)
        x = x.view(b, nh, nh, ws, ws, c).transpose(2, 3).reshape(b, ph, ph, c)
        if pad:
            x = x[:, :h, :w].contiguous()
        x = res_x + x.view(b, l, c)
        x = x.transpose(1, 2).reshape(b, c, h, w)
        x = self.local_conv(x)
        x = x.view(b, c, l).transpose(1, 2)
        return x + self.mlp(x)


class BasicLayer(nn.Module):
    def __init__(self, dim, res, depth, heads, ws, out_dim, downsample):
        super().__init__()
        self.blocks = nn.ModuleList([TinyViTBlock(dim, res, heads, ws) for _ in range(depth)])
        self.downsample = PatchMerging(res, dim, out_dim) if downsample else None

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return self.downsample(x) if self.downsample is not None else x


class LayerNorm2d(nn.Module):
    def __init__(self, c, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.eps = eps

    def forward(self, x):
        u = x.mean(1, keepdim=True)
        s = (x - u).pow(2).mean(1, keepdim=True)
        x = (x - u) / torch.sqrt(s + self.eps)
        return self.weight[:, None, None] * x + self.bias[:, None, None]


class TinyViT(nn.Module):
    """MobileSAM image encoder; `num_classes` head kept so the parameter count matches the checkpoint."""

    def __init__(self, img_size=IMG_SIZE, dims=(64, 128, 160, 320), depths=(2, 2, 6, 2), heads=(2, 4, 5, 10),
                 windows=(7, 7, 14, 7), num_classes=1000):
        super().__init__()
        self.img_size = img_size
        self.patch_embed = PatchEmbed(3, dims[0])
        pr = img_size // 4
        self.layers = nn.ModuleList()
        for i in range(4):
            res = pr // (2 ** (i - 1 if i == 3 else i))
            out_dim = dims[min(i + 1, 3)]
            if i == 0:
                self.layers.append(ConvLayer(dims[0], res, depths[0], out_dim, 4.0))
            else:
                self.layers.append(BasicLayer(dims[i], res, depths[i], heads[i], windows[i], out_dim, i < 3))
        self.norm_head = nn.LayerNorm(dims[-1])  # unused by forward (kept for the state dict)
        self.head = nn.Linear(dims[-1], num_classes)  # unused by forward
        self.neck = nn.Sequential(
            nn.Conv2d(dims[-1], 256, 1, bias=False), LayerNorm2d(256),
            nn.Conv2d(256, 256, 3, padding=1, bias=False), LayerNorm2d(256))
        self.final_res = pr // 4

    def forward(self, x, taps: dict | None = None):
        x = self.patch_embed(x)
        if taps is not None:
            taps["patch_embed"] = x
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if taps is not None:
                taps[f"layer{i}"] = x
        b, _, c = x.shape
        x = x.view(b, self.final_res, self.final_res, c).permute(0, 3, 1, 2)
        return self.neck(x)


class EncoderWithPreprocess(nn.Module):
    """`use_preprocess=True` export wrapper (reference script/export_models.py:26; Appendix A.1):
    input (H, W, 3) f32 RGB 0..255 -> normalise -> CHW -> zero-pad to 1024^2 -> TinyViT."""

    def __init__(self, enc: TinyViT):
        super().__init__()
        self.enc = enc
        self.register_buffer("mean", torch.tensor(PIXEL_MEAN).view(1, 1, 3), persistent=False)
        self.register_buffer("std", torch.tensor(PIXEL_STD).view(1, 1, 3), persistent=False)

    def preprocess(self, img_hwc):
        x = (img_hwc - self.mean) / self.std
        x = x.permute(2, 0, 1)
        h, w = x.shape[-2:]
        x = F.pad(x, (0, self.enc.img_size - w, 0, self.enc.img_size - h))
        return x[None]

    def forward(self, img_hwc, taps=None):
        return self.enc(self.preprocess(img_hwc), taps)


# --------------------------------------------------------------------------------------------
# Prompt encoder + mask decoder (standard SAM, transformer_dim 256): Appendix A.4
# --------------------------------------------------------------------------------------------
class PositionEmbeddingRandom(nn.Module):
    def __init__(self, num_pos_feats=128):
        super().__init__()
        self.register_buffer("positional_encoding_gaussian_matrix", torch.randn(2, num_pos_feats))

    def _pe_encoding(self, coords):
        coords = 2 * coords - 1
        coords = coords @ self.positional_encoding_gaussian_matrix
        coords = 2 * math.pi * coords
        return torch.cat([torch.sin(coords), torch.cos(coords)], dim=-1)

    def forward(self, size):
        h, w = size
        grid = torch.ones(h, w, dtype=torch.float32)
        y = (grid.cumsum(0) - 0.5) / h
        x = (grid.cumsum(1) - 0.5) / w
        return self._pe_encoding(torch.stack([x, y], dim=-1)).permute(2, 0, 1)


class PromptEncoder(nn.Module):
    def __init__(self, dim=256, emb_size=64, mask_in=16):
        super().__init__()
        self.emb_size = emb_size
        self.pe_layer = PositionEmbeddingRandom(dim // 2)
        self.point_embeddings = nn.ModuleList([nn.Embedding(1, dim) for _ in range(4)])
        self.not_a_point_embed = nn.Embedding(1, dim)
        self.no_mask_embed = nn.Embedding(1, dim)
        self.mask_downscaling = nn.Sequential(
            nn.Conv2d(1, mask_in // 4, 2, 2), LayerNorm2d(mask_in // 4), nn.GELU(),
            nn.Conv2d(mask_in // 4, mask_in, 2, 2), LayerNorm2d(mask_in), nn.GELU(),
            nn.Conv2d(mask_in, dim, 1))

    def get_dense_pe(self):
        return self.pe_layer((self.emb_size, self.emb_size))[None]


class SamAttention(nn.Module):
    def __init__(self, dim, heads, downsample=1):
        super().__init__()
        self.internal = dim // downsample
        self.heads = heads
        self.q_proj = nn.Linear(dim, self.internal)
        self.k_proj = nn.Linear(dim, self.internal)
        self.v_proj = nn.Linear(dim, self.internal)
        self.out_proj = nn.Linear(self.internal, dim)

    def _split(self, x):
        b, n, c = x.shape
        return x.reshape(b, n, self.heads, c // self.heads).transpose(1, 2)

    def forward(self, q, k, v):
        q, k, v = self._split(self.q_proj(q)), self._split(self.k_proj(k)), self._split(self.v_proj(v))
        d = q.shape[-1]
        attn = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(d), dim=-1)
        out = (attn @ v).transpose(1, 2)
        b, n, h, d = out.shape
        return self.out_proj(out.reshape(b, n, h * d))


class MLPBlock(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.lin1 = nn.Linear(dim, hidden)
        self.lin2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.lin2(F.relu(self.lin1(x)))


class TwoWayAttentionBlock(nn.Module):
    def __init__(self, dim, heads, mlp_dim, skip_first_layer_pe):
        super().__init__()
        self.self_attn = SamAttention(dim, heads)
        self.norm1 = nn.LayerNorm(dim)
        self.cross_attn_token_to_image = SamAttention(dim, heads, 2)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = MLPBlock(dim, mlp_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.norm4 = nn.LayerNorm(dim)
        self.cross_attn_image_to_token = SamAttention(dim, heads, 2)
        self.skip_first_layer_pe = skip_first_layer_pe

    def forward(self, queries, keys, query_pe, key_pe):
        if self.skip_first_layer_pe:
            queries = self.self_attn(queries, queries, queries)
        else:
            q = queries + query_pe
            queries = queries + self.self_attn(q, q, queries)
        queries = self.norm1(queries)
        q = queries + query_pe
        k = keys + key_pe
        queries = self.norm2(queries + self.cross_attn_token_to_image(q, k, keys))
        queries = self.norm3(queries + self.mlp(queries))
        q = queries + query_pe
        k = keys + key_pe
        keys = self.norm4(keys + self.cross_attn_image_to_token(k, q, queries))
        return queries, keys


class TwoWayTransformer(nn.Module):
    def __init__(self, depth=2, dim=256, heads=8, mlp_dim=2048):
        super().__init__()
        self.layers = nn.ModuleList([TwoWayAttentionBlock(dim, heads, mlp_dim, i == 0) for i in range(depth)])
        self.final_attn_token_to_image = SamAttention(dim, heads, 2)
        self.norm_final_attn = nn.LayerNorm(dim)

    def forward(self, image_embedding, image_pe, point_embedding):
        image_embedding = image_embedding.flatten(2).permute(0, 2, 1)
        image_pe = image_pe.flatten(2).permute(0, 2, 1)
        queries, keys = point_embedding, image_embedding
        for layer in self.layers:
            queries, keys = layer(queries, keys, point_embedding, image_pe)
        q = queries + point_embedding
        k = keys + image_pe
        queries = self.norm_final_attn(queries + self.final_attn_token_to_image(q, k, keys))
        return queries, keys


class MLP(nn.Module):
    def __init__(self, din, hid, dout, n):
        super().__init__()
        dims = [din] + [hid] * (n - 1) + [dout]
        self.layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])

    def forward(self, x):
        for i, l in enumerate(self.layers):
            x = F.relu(l(x)) if i < len(self.layers) - 1 else l(x)
        return x


class MaskDecoder(nn.Module):
    def __init__(self, dim=256, num_multimask=3):
        super().__init__()
        self.transformer = TwoWayTransformer()
        self.num_mask_tokens = num_multimask + 1
        self.iou_token = nn.Embedding(1, dim)
        self.mask_tokens = nn.Embedding(self.num_mask_tokens, dim)
        self.output_upscaling = nn.Sequential(
            nn.ConvTranspose2d(dim, dim // 4, 2, 2), LayerNorm2d(dim // 4), nn.GELU(),
            nn.ConvTranspose2d(dim // 4, dim // 8, 2, 2), nn.GELU())
        self.output_hypernetworks_mlps = nn.ModuleList([MLP(dim, dim, dim // 8, 3) for _ in range(self.num_mask_tokens)])
        self.iou_prediction_head = MLP(dim, 256, self.num_mask_tokens, 3)

    def predict_masks(self, image_embeddings, image_pe, sparse, dense):
        out_tokens = torch.cat([self.iou_token.weight, self.mask_tokens.weight], dim=0)
        out_tokens = out_tokens[None].expand(sparse.size(0), -1, -1)
        tokens = torch.cat((out_tokens, sparse), dim=1)
        src = torch.repeat_interleave(image_embeddings, tokens.shape[0], dim=0) + dense
        pos = torch.repeat_interleave(image_pe, tokens.shape[0], dim=0)
        b, c, h, w = src.shape
        hs, src = self.transformer(src, pos, tokens)
        iou_tok = hs[:, 0, :]
        mask_toks = hs[:, 1:1 + self.num_mask_tokens, :]
        src = src.transpose(1, 2).view(b, c, h, w)
        up = self.output_upscaling(src)
        hyper = torch.stack([self.output_hypernetworks_mlps[i](mask_toks[:, i, :])
                             for i in range(self.num_mask_tokens)], dim=1)
        b, c, h, w = up.shape
        masks = (hyper @ up.view(b, c, h * w)).view(b, -1, h, w)
        return masks, self.iou_prediction_head(iou_tok)


class MobileSam(nn.Module):
    """Container with the checkpoint's top-level names (image_encoder / prompt_encoder / mask_decoder)."""

    def __init__(self):
        super().__init__()
        self.image_encoder = TinyViT()
        self.prompt_encoder = PromptEncoder()
        self.mask_decoder = MaskDecoder()


class SamOnnxDecoder(nn.Module):
    """The decoder export wrapper (segment-anything `SamOnnxModel`; Appendix A.5), i.e. what
    `sam_mask_decoder_{single,multi}.onnx` compute for the six inputs named at
    reference src/segmentation.cpp:19-22."""

    def __init__(self, sam: MobileSam, return_single_mask: bool):
        super().__init__()
        self.sam = sam
        self.single = return_single_mask
        self.img_size = IMG_SIZE

    def embed_points(self, coords, labels):
        pe = self.sam.prompt_encoder
        c = (coords + 0.5) / self.img_size
        emb = pe.pe_layer._pe_encoding(c)
        lab = labels.unsqueeze(-1).expand_as(emb)
        emb = emb * (lab != -1)
        emb = emb + pe.not_a_point_embed.weight * (lab == -1)
        for i in range(4):
            emb = emb + pe.point_embeddings[i].weight * (lab == i)
        return emb

    def embed_masks(self, mask_input, has_mask):
        pe = self.sam.prompt_encoder
        e = has_mask * pe.mask_downscaling(mask_input)
        return e + (1 - has_mask) * pe.no_mask_embed.weight.reshape(1, -1, 1, 1)

    @staticmethod
    def prepadded_size(orig_hw: torch.Tensor, longest: int):
        orig_hw = orig_hw.to(torch.float32)
        scale = longest / torch.max(orig_hw)
        return torch.floor(scale * orig_hw + 0.5).to(torch.int64)

    def postprocess(self, masks, orig_hw):
        masks = F.interpolate(masks, size=(self.img_size, self.img_size), mode="bilinear", align_corners=False)
        pp = self.prepadded_size(orig_hw, self.img_size)
        masks = masks[..., : int(pp[0]), : int(pp[1])]
        h, w = int(orig_hw[0]), int(orig_hw[1])
        return F.interpolate(masks, size=(h, w), mode="bilinear", align_corners=False)

    def select(self, masks, iou, num_points):
        rw = torch.tensor([[1000.0] + [0.0] * (masks.shape[1] - 1)])
        score = iou + (num_points - 2.5) * rw
        best = torch.argmax(score, dim=1)
        ar = torch.arange(masks.shape[0])
        return masks[ar, best][:, None], iou[ar, best][:, None]

    def low_res(self, image_embeddings, point_coords, point_labels):
        sparse = self.embed_points(point_coords, point_labels)
        dense = self.embed_masks(torch.zeros(1, 1, 256, 256), torch.zeros(1))
        masks, iou = self.sam.mask_decoder.predict_masks(
            image_embeddings, self.sam.prompt_encoder.get_dense_pe(), sparse, dense)
        if self.single:
            masks, iou = self.select(masks, iou, point_coords.shape[1])
        return masks, iou

    def forward(self, image_embeddings, point_coords, point_labels, mask_input, has_mask_input, orig_im_size):
        sparse = self.embed_points(point_coords, point_labels)
        dense = self.embed_masks(mask_input, has_mask_input)
        masks, iou = self.sam.mask_decoder.predict_masks(
            image_embeddings, self.sam.prompt_encoder.get_dense_pe(), sparse, dense)
        if self.single:
            masks, iou = self.select(masks, iou, point_coords.shape[1])
        return self.postprocess(masks, orig_im_size), iou, masks


# --------------------------------------------------------------------------------------------
# Weights: the oracle loads the SAME tensors the engine loads (state-dict names of SURVEY A.7)
# --------------------------------------------------------------------------------------------
def load_numpy_state(sam: MobileSam, tensors) -> MobileSam:
    """tensors: {state-dict name: float32 ndarray} (dlimgedit_b200.weights_io container contents).  Strict:
    every floating-point entry of the model's state dict must be present with the right shape."""
    sd = sam.state_dict()
    with torch.no_grad():
        for k, v in sd.items():
            if not v.dtype.is_floating_point:
                continue  # BatchNorm num_batches_tracked
            if k not in tensors:
                raise KeyError(f"weight container lacks {k}")
            t = torch.from_numpy(tensors[k])
            if tuple(t.shape) != tuple(v.shape):
                raise ValueError(f"{k}: container shape {tuple(t.shape)} != model shape {tuple(v.shape)}")
            v.copy_(t)
    extra = set(tensors) - set(sd)
    if extra:
        raise KeyError(f"weight container has unknown tensors: {sorted(extra)[:5]}")
    sam.eval()
    return sam


def build_synthetic(seed: int = 0, stress: bool = False) -> MobileSam:
    """Oracle model carrying the seeded synthetic weights of dlimgedit_b200.synthetic_weights (no checkpoint
    exists offline); stress = heavy-tailed norm scales and strong residual branches (see make_state_dict)."""
    from dlimgedit_b200 import synthetic_weights
    return load_numpy_state(MobileSam(), synthetic_weights.make_state_dict(seed, stress))


def count_learnable(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())
