"""Self-consistency gates for the PyTorch oracle of the MobileSAM graphs (CPU only).

PARITY UNPINNED against the real ONNX graphs (not available offline).  What is checked here:
  * learnable-parameter count and split match the published MobileSAM / SAM sizes,
  * everything but the TinyViT trunk agrees with the independent restatement in `transformers.models.sam` on copied weights:
    mask decoder, prompt encoder + dense positional grid, mask post-processing + threshold, resized-extent rule,
    normalise + pad preprocessing,
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


def _hf_prompt_encoder(sam):
    from transformers.models.sam.configuration_sam import SamConfig
    from transformers.models.sam.modeling_sam import SamPromptEncoder
    hf = SamPromptEncoder(SamConfig()).eval()
    pe = sam.prompt_encoder
    with torch.no_grad():
        hf.shared_embedding.positional_embedding.copy_(pe.pe_layer.positional_encoding_gaussian_matrix)
        for i in range(4):
            hf.point_embed[i].weight.copy_(pe.point_embeddings[i].weight)
        hf.not_a_point_embed.weight.copy_(pe.not_a_point_embed.weight)
        hf.no_mask_embed.weight.copy_(pe.no_mask_embed.weight)
    return hf


def test_prompt_encoder_matches_transformers_sam(oracle_sam):
    """The prompt assembly of the exported decoder graph (SamOnnxModel._embed_points: point / padding point / box corners as
    labelled points) against `transformers.models.sam.SamPromptEncoder`, an independent restatement of SAM's prompt
    encoder with the same random-Fourier positional encoding: a point prompt (label 1 + the padding point the reference appends,
    segmentation.cpp:134-143), a box prompt (corner labels 2 / 3, segmentation.cpp:144-152), the no-mask dense embedding and
    the dense positional grid."""
    sam = oracle_sam
    hf = _hf_prompt_encoder(sam)
    dec = R.SamOnnxDecoder(sam, return_single_mask=True)
    with torch.no_grad():
        # point (276, 411) on the 1024 grid + padding point at the transformed origin, labels 1 / -1
        coords = torch.tensor([[[276.0, 411.0], [0.0, 0.0]]])
        labels = torch.tensor([[1.0, -1.0]])
        ours = dec.embed_points(coords, labels)
        theirs, dense_hf = hf(torch.tensor([[[[276.0, 411.0]]]]), torch.tensor([[[1]]]), None, None)  # pads by itself
        assert theirs.shape == (1, 1, 2, 256)
        assert torch.allclose(ours, theirs[:, 0], atol=1e-6, rtol=1e-5)
        # box (102, 63) - (287, 188): two corner points with labels 2 / 3 and no padding point
        box = torch.tensor([[[102.0, 63.0], [287.0, 188.0]]])
        ours_b = dec.embed_points(box, torch.tensor([[2.0, 3.0]]))
        theirs_b, _ = hf(None, None, torch.tensor([[[102.0, 63.0, 287.0, 188.0]]]), None)
        assert torch.allclose(ours_b, theirs_b[:, 0], atol=1e-6, rtol=1e-5)
        # the constant mask input of the reference (segmentation.cpp:43-45: zeros, has_mask_input = 0)
        dense = dec.embed_masks(torch.zeros(1, 1, 256, 256), torch.zeros(1))
        assert torch.equal(dense.expand(1, 256, 64, 64), dense_hf)
        # dense positional encoding of the 64 x 64 embedding grid
        from transformers.models.sam.configuration_sam import SamConfig
        from transformers.models.sam.modeling_sam import SamModel
        grid_hf = SamModel.get_image_wide_positional_embeddings.__get__(
            type("M", (), {"config": SamConfig(), "shared_image_embedding": hf.shared_embedding})())()
        assert torch.allclose(sam.prompt_encoder.get_dense_pe(), grid_hf, atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("h,w", [(1200, 1800), (683, 1024), (1024, 600), (2160, 3840), (37, 53)])
def test_mask_postprocessing_matches_transformers_sam(oracle_sam, h, w):
    """The exported decoder's mask post-processing (SamOnnxModel.mask_postprocessing: 256 -> 1024 bilinear, crop to the resized
    extent, bilinear to the original extent) and the reference's `> 0` threshold (segmentation.cpp:
    108-116) against `transformers`' independent `SamImageProcessor.post_process_masks`, including the round-half-up rule for
    the resized extent (ResizeLongestSide, segmentation.cpp:60-70)."""
    from transformers.models.sam.image_processing_sam import SamImageProcessor
    from oracle import prepost as P
    dec = R.SamOnnxDecoder(oracle_sam, True)
    g = torch.Generator().manual_seed(h * 7 + w)
    low = torch.randn(1, 1, 256, 256, generator=g) * 3
    ours = dec.postprocess(low, torch.tensor([float(h), float(w)]))
    pp = dec.prepadded_size(torch.tensor([float(h), float(w)]), 1024)
    _need, rw, rh, _scale = P.resize_longest_side(w, h)
    assert (int(pp[0]), int(pp[1])) == (rh, rw)          # the oracle's two statements of the rule agree
    proc = SamImageProcessor()
    theirs = proc.post_process_masks([low], [[h, w]], [[int(pp[0]), int(pp[1])]], binarize=False)[0]
    assert theirs.shape == ours.shape == (1, 1, h, w)
    assert torch.allclose(ours, theirs, atol=1e-5, rtol=1e-5)
    assert torch.equal(ours > 0, proc.post_process_masks([low], [[h, w]], [[int(pp[0]), int(pp[1])]], binarize=True)[0])
    # and HF's own statement of the resized extent (its pre-processing resize) is the same round-half-up
    out_hw = proc._get_preprocess_shape((h, w), 1024)
    assert tuple(out_hw) == (rh, rw)


def test_image_tensor_and_preprocessing_match_transformers_sam(oracle_sam):
    """create_image_tensor (reference segmentation.cpp:81-106: channel map to RGB floats 0..255) followed by the exported
    encoder's own preprocessing ((x - mean) / std with SAM's constants, zero padding to 1024 x 1024 AFTER normalisation, planar
    layout) against `transformers`' SamImageProcessor with its resize switched off (that one resamples with PIL, the
    reference with stb's Mitchell filter: not comparable)."""
    from transformers.models.sam.image_processing_sam import SamImageProcessor
    from oracle import prepost as P
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (683, 1024, 3), dtype=np.uint8)
    enc = R.EncoderWithPreprocess(oracle_sam.image_encoder)
    t = P.create_image_tensor(rgb, 3)                          # Channels::rgb
    assert t.shape == (683, 1024, 3) and np.array_equal(t, rgb.astype(np.float32))
    ours = enc.preprocess(torch.from_numpy(t)).numpy()
    theirs = SamImageProcessor(do_resize=False)(images=rgb, return_tensors="np")["pixel_values"]
    assert theirs.shape == ours.shape == (1, 3, 1024, 1024)
    assert np.allclose(ours, theirs, atol=2e-6, rtol=1e-6)
    assert np.all(ours[:, :, 683:, :] == 0) and np.all(theirs[:, :, 683:, :] == 0)
    bgra = np.dstack([rgb[..., ::-1], rng.integers(0, 256, (683, 1024), dtype=np.uint8)])
    assert np.array_equal(P.create_image_tensor(bgra, 5), t)   # Channels::bgra (dlimgedit.hpp:29): same planes, alpha dropped


@pytest.mark.parametrize("ws", [7, 14])
def test_attention_bias_index_table_matches_levit(ws):
    """TinyViT's attention takes its relative-position bias table from LeViT (first-seen ordering of (|dy|, |dx|) offsets over
    itertools.product pairs); `transformers`' LevitAttention builds the same table independently."""
    from transformers.models.levit.modeling_levit import LevitAttention
    att = LevitAttention(hidden_sizes=32, key_dim=8, num_attention_heads=2, attention_ratio=1, resolution=ws)
    idx, n_off = R.attention_bias_idxs(ws)
    assert n_off == att.attention_biases.shape[1] == ws * ws
    assert torch.equal(idx, att.attention_bias_idxs)
