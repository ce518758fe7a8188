"""Self-consistency gates for the PyTorch oracle of the MobileSAM graphs (CPU only).

PARITY UNPINNED against the real ONNX graphs (not available offline).  What is checked here:
  * learnable-parameter count and split match the published MobileSAM / SAM sizes,
  * the decoder half agrees with the independent restatement in `transformers.models.sam`,
  * the ONNX wrapper's selection / post-processing rules.
"""
import numpy as np
import pytest
import torch

from oracle import mobile_sam_ref as R


def test_parameter_count(oracle_sam):
    sam = oracle_sam
    assert R.count_learnable(sam) == 10_130_092
    enc = sam.image_encoder
    unused = R.count_learnable(enc.norm_head) + R.count_learnable(enc.head)
    assert unused == 321_640
    assert R.count_learnable(enc) - unused == 5_743_892
    assert R.count_learnable(enc.neck) == 672_768
    assert R.count_learnable(sam.prompt_encoder) == 6_220
    md = sam.mask_decoder
    assert R.count_learnable(md) == 4_058_340
    assert R.count_learnable(md.transformer) == 3_291_264
    assert R.count_learnable(md.output_upscaling) == 73_952
    assert R.count_learnable(md.output_hypernetworks_mlps) == 559_232
    assert R.count_learnable(md.iou_prediction_head) == 132_612


def test_attention_bias_index_table():
    idx, n_off = R.attention_bias_idxs(7)
    assert idx.shape == (49, 49) and n_off == 49
    assert idx[0, 0] == 0 and idx[0, 1] == 1 and idx[1, 0] == 1  # symmetric in |d|
    idx14, n14 = R.attention_bias_idxs(14)
    assert idx14.shape == (196, 196) and n14 == 196


def test_encoder_shapes_and_window_padding(oracle_sam):
    enc = R.EncoderWithPreprocess(oracle_sam.image_encoder)
    img = torch.zeros(40, 64, 3)
    x = enc.preprocess(img)
    assert x.shape == (1, 3, 1024, 1024)
    # zero padding happens AFTER normalisation: padded area is exactly 0, image area is -mean/std
    assert float(x[0, 0, 100, 100]) == 0.0
    assert abs(float(x[0, 0, 0, 0]) + 123.675 / 58.395) < 1e-6


def _hf_decoder(sam):
    from transformers.models.sam.configuration_sam import SamMaskDecoderConfig
    from transformers.models.sam.modeling_sam import SamMaskDecoder
    cfg = SamMaskDecoderConfig(layer_norm_eps=1e-5)  # original SAM uses nn.LayerNorm's default eps
    cfg._attn_implementation = "eager"
    hf = SamMaskDecoder(cfg).eval()
    sd = {}
    for k, v in sam.mask_decoder.state_dict().items():
        k2 = k.replace(".norm1.", ".layer_norm1.").replace(".norm2.", ".layer_norm2.").replace(".norm3.", ".layer_norm3.")
        k2 = k2.replace(".norm4.", ".layer_norm4.").replace("norm_final_attn", "layer_norm_final_attn")
        k2 = k2.replace("output_upscaling.0.", "upscale_conv1.").replace("output_upscaling.1.", "upscale_layer_norm.")
        k2 = k2.replace("output_upscaling.3.", "upscale_conv2.")
        for head in ("output_hypernetworks_mlps.0", "output_hypernetworks_mlps.1", "output_hypernetworks_mlps.2",
                     "output_hypernetworks_mlps.3", "iou_prediction_head"):
            if k2.startswith(head + ".layers."):
                rest = k2[len(head) + len(".layers."):]
                i, tail = rest.split(".", 1)
                k2 = head + {"0": ".proj_in.", "1": ".layers.0.", "2": ".proj_out."}[i] + tail
        sd[k2] = v
    missing, unexpected = hf.load_state_dict(sd, strict=True)
    return hf


def test_decoder_matches_transformers_sam(oracle_sam):
    sam = oracle_sam
    hf = _hf_decoder(sam)
    g = torch.Generator().manual_seed(5)
    emb = torch.randn(1, 256, 64, 64, generator=g)
    dec = R.SamOnnxDecoder(sam, return_single_mask=False)
    coords = torch.tensor([[[300.0, 410.0], [0.0, 0.0]]])
    labels = torch.tensor([[1.0, -1.0]])
    with torch.no_grad():
        sparse = dec.embed_points(coords, labels)
        dense = dec.embed_masks(torch.zeros(1, 1, 256, 256), torch.zeros(1))
        pe = sam.prompt_encoder.get_dense_pe()
        masks, iou = sam.mask_decoder.predict_masks(emb, pe, sparse, dense)
        m_hf, iou_hf = hf(emb, pe, sparse[:, None], dense, multimask_output=True)
    assert masks.shape == (1, 4, 256, 256)
    assert torch.allclose(masks[:, 1:], m_hf[:, 0], atol=2e-4, rtol=1e-4)
    assert torch.allclose(iou[:, 1:], iou_hf[:, 0], atol=1e-5, rtol=1e-4)


def test_single_mask_selection_rule(oracle_sam):
    dec = R.SamOnnxDecoder(oracle_sam, True)
    masks = torch.arange(4.0).view(1, 4, 1, 1).expand(1, 4, 2, 2)
    # token 0 is pushed down by (2 - 2.5) * 1000 even when its IoU is the largest
    m, s = dec.select(masks, torch.tensor([[0.99, 0.2, 0.7, 0.3]]), 2)
    assert float(m[0, 0, 0, 0]) == 2.0 and float(s) == pytest.approx(0.7)


def test_postprocess_crops_to_resized_extent(oracle_sam):
    dec = R.SamOnnxDecoder(oracle_sam, True)
    assert dec.prepadded_size(torch.tensor([1200.0, 1800.0]), 1024).tolist() == [683, 1024]
    low = torch.zeros(1, 1, 256, 256)
    low[..., :171, :] = 1.0   # rows that map inside the 683-row valid area
    low[..., 171:, :] = -1.0  # padding area must not leak into the output
    out = dec.postprocess(low, torch.tensor([1200.0, 1800.0]))
    assert out.shape == (1, 1, 1200, 1800)
    assert float((out > 0).float().mean()) > 0.99
