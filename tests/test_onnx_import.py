"""SURVEY 8(f2) real-weight ingestion: the ONNX initializer reader (dlimgedit_b200/onnx_import.py) and the checkpoint
converter (tools/convert_checkpoint.py).  No real .onnx / .pt exists offline, so the reader runs on graphs written by
tests/onnx_emit.py from the oracle's state dict with the exporter's conventions (folded BatchNorm, transposed anonymous
MatMul weights, decomposed LayerNorm2d, constant-folded bias gathers), and the result must reproduce the model.  CPU only."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from dlimgedit_b200 import onnx_import, synthetic_weights, weights_io
from oracle.mobile_sam_ref import EncoderWithPreprocess, MobileSam, SamOnnxDecoder, load_numpy_state

import onnx_emit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def onnx_dir(tmp_path_factory, oracle_sam):
    d = tmp_path_factory.mktemp("onnx")
    sd = synthetic_weights.make_state_dict(0)
    dense_pe = oracle_sam.prompt_encoder.get_dense_pe().numpy()
    (d / onnx_import.ENCODER_ONNX).write_bytes(onnx_emit.encoder_onnx(sd))
    for name in onnx_import.DECODER_ONNX:
        (d / name).write_bytes(onnx_emit.decoder_onnx(sd, dense_pe))
    return str(d), sd


def _fold(sd, p):
    scale = sd[p + ".bn.weight"] / np.sqrt(sd[p + ".bn.running_var"] + 1e-5)
    return sd[p + ".c.weight"] * scale[:, None, None, None], sd[p + ".bn.bias"] - sd[p + ".bn.running_mean"] * scale


def test_protobuf_reader_roundtrip():
    e = onnx_emit.Emitter()
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4) - 7.5
    i64 = np.array([[-1, 2], [1 << 40, -(1 << 40)]], np.int64)
    e.init("a_raw", a)            # raw_data
    e.init("pad", np.zeros(3))
    e.init("a_typed", a)          # float_data (every third tensor is written with the typed repeated field)
    e.init("i_raw", i64, dtype=np.int64)
    e.init("pad2", np.zeros(1))
    e.init("i_typed", i64, dtype=np.int64)
    x = e.node("Conv", ["x", "a_raw"], {"group": 3, "kernel_shape": [3, 3], "alpha": 0.5})
    c = e.scalar(2.5)
    g = onnx_import.parse_model(e.model(["x"], [x]))
    assert g.inputs == ["x"] and g.outputs == [x]
    for k in ("a_raw", "a_typed"):
        assert g.initializers[k].dtype == np.float32 and np.array_equal(g.initializers[k], a)
    for k in ("i_raw", "i_typed"):
        assert g.initializers[k].dtype == np.int64 and np.array_equal(g.initializers[k], i64)
    conv = [n for n in g.nodes if n.op == "Conv"][0]
    assert conv.attrs["group"] == 3 and conv.attrs["kernel_shape"] == [3, 3] and conv.attrs["alpha"] == 0.5
    assert float(g.initializers[c]) == 2.5  # Constant node -> initializer
    with pytest.raises(ValueError):
        onnx_import.parse_model(b"\x08\x08")  # no graph


def test_onnx_weights_come_back_under_state_dict_names(onnx_dir):
    d, sd = onnx_dir
    got = onnx_import.state_from_onnx_dir(d)
    unused = {k for k in sd if k.startswith("image_encoder.norm_head") or k.startswith("image_encoder.head")}
    assert set(got) == set(sd) - unused
    for k, v in got.items():
        if ".bn." in k or k.endswith(".c.weight"):
            continue
        assert v.shape == sd[k].shape and np.array_equal(v, sd[k]), k
    for p in sorted({k[: -len(".c.weight")] for k in sd if k.endswith(".c.weight")}):  # folded Conv + BN pairs
        w_ref, b_ref = _fold(sd, p)
        w_got, b_got = _fold(got, p)
        assert np.allclose(w_got, w_ref, rtol=1e-6, atol=1e-7) and np.allclose(b_got, b_ref, rtol=1e-6, atol=1e-7), p


def test_imported_weights_reproduce_the_model(onnx_dir, oracle_sam, tmp_path):
    """tools/convert_checkpoint.py --onnx-dir -> container -> a model that computes what the original does."""
    d, sd = onnx_dir
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "convert_checkpoint.py"), "--onnx-dir", d, str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    tensors = weights_io.load(str(tmp_path / "segmentation" / weights_io.WEIGHT_FILE_NAME))
    for k in sd:  # tensors the graphs never use (classification head) are not in the files: zero them for the strict loader
        tensors.setdefault(k, np.zeros_like(sd[k]))
    sam2 = load_numpy_state(MobileSam(), tensors)
    torch.manual_seed(0)
    img = torch.rand(96, 128, 3) * 255
    with torch.no_grad():
        e1 = EncoderWithPreprocess(oracle_sam.image_encoder)(img)
        e2 = EncoderWithPreprocess(sam2.image_encoder)(img)
        assert torch.allclose(e1, e2, atol=2e-4, rtol=1e-4), float((e1 - e2).abs().max())
        coords, labels = torch.tensor([[[300.0, 410.0], [0.0, 0.0]]]), torch.tensor([[1.0, -1.0]])
        l1, i1 = SamOnnxDecoder(oracle_sam, False).low_res(e1, coords, labels)
        l2, i2 = SamOnnxDecoder(sam2, False).low_res(e1, coords, labels)
    assert torch.equal(l1, l2) and torch.equal(i1, i2)  # decoder weights come back bit for bit


def test_checkpoint_converter_roundtrip(tmp_path):
    """torch.save(state_dict) -> tools/convert_checkpoint.py -> the container the synthetic generator writes directly."""
    sd = synthetic_weights.make_state_dict(3)
    ckpt = {k: torch.from_numpy(v) for k, v in sd.items()}
    ckpt["image_encoder.layers.1.blocks.0.attn.attention_bias_idxs"] = torch.zeros(49, 49, dtype=torch.long)  # integer buffers are dropped
    ckpt["image_encoder.patch_embed.seq.0.bn.num_batches_tracked"] = torch.tensor(7)
    torch.save({"model": ckpt}, tmp_path / "mobile_sam.pt")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "convert_checkpoint.py"), str(tmp_path / "mobile_sam.pt"), str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    direct = tmp_path / "direct"
    synthetic_weights.write_model_dir(str(direct), seed=3)
    a = (tmp_path / "segmentation" / weights_io.WEIGHT_FILE_NAME).read_bytes()
    b = (direct / "segmentation" / weights_io.WEIGHT_FILE_NAME).read_bytes()
    assert a == b
    bad = tmp_path / "other.pt"
    torch.save({"some.weight": torch.zeros(3)}, bad)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "convert_checkpoint.py"), str(bad), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode != 0 and "does not look like a MobileSAM" in r.stderr
