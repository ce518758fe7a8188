"""Test support: writes ONNX files shaped like the reference's three MobileSAM graphs from a MobileSAM state dict, with
the conventions of torch.onnx.export (opset 17, constant folding on) that dlimgedit_b200/onnx_import.py has to undo:

  * Conv2d_BN pairs folded into one Conv with anonymous `onnx::Conv_N` weight / bias,
  * nn.Linear as MatMul with the TRANSPOSED weight under an anonymous `onnx::MatMul_N` name + Add(bias), or Gemm(transB=1),
  * nn.LayerNorm as LayerNormalization, LayerNorm2d decomposed into ReduceMean / Sub / Pow / ... / Mul(weight) / Add(bias),
  * the relative-position bias either as Gather(attention_biases, idxs) or constant-folded to a dense (1, heads, n, n) Add,
  * exact-erf GELU decomposed into Div / Erf / Add / Mul / Mul with scalar constants, Constant nodes, int64 shape tensors,
  * [iou_token; mask_tokens] folded into one (5, 256) initializer, the dense positional encoding folded to a constant.

Only the protobuf subset the importer reads is written (ModelProto.graph, GraphProto.node / initializer / input / output,
NodeProto, AttributeProto i / f / t / ints, TensorProto raw_data or typed fields).  The graphs are NOT meant to be executed:
node inputs are chained plausibly, which is all a weight importer looks at.  No real .onnx file exists offline."""
from __future__ import annotations

import struct
from typing import Dict, List

import numpy as np

from dlimgedit_b200.onnx_import import BN_EPS, DEPTHS, DIMS, HEADS, WINDOWS, attention_bias_idxs


def _varint(x: int) -> bytes:
    if x < 0:
        x += 1 << 64
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _key(field: int, wt: int) -> bytes:
    return _varint((field << 3) | wt)


def _ld(field: int, payload: bytes) -> bytes:
    return _key(field, 2) + _varint(len(payload)) + payload


def tensor_proto(name: str, arr: np.ndarray, raw: bool = True) -> bytes:
    arr = np.asarray(arr)
    code = {np.dtype(np.float32): 1, np.dtype(np.int64): 7, np.dtype(np.float64): 11}[arr.dtype]
    out = b"".join(_key(1, 0) + _varint(int(d)) for d in arr.shape)
    out += _key(2, 0) + _varint(code)
    if raw:
        out += _ld(9, arr.astype(arr.dtype.newbyteorder("<")).tobytes())
    elif code == 1:
        out += _ld(4, arr.astype("<f4").tobytes())  # packed float_data
    elif code == 7:
        out += _ld(7, b"".join(_varint(int(v)) for v in arr.reshape(-1)))
    else:
        out += _ld(10, arr.astype("<f8").tobytes())
    out += _ld(8, name.encode())
    return out


class Emitter:
    def __init__(self):
        self.nodes: List[bytes] = []
        self.inits: List[bytes] = []
        self.counter = 100
        self.flip = 0

    def anon(self, kind: str) -> str:
        self.counter += 7
        return f"onnx::{kind}_{self.counter}"

    def init(self, name: str, arr, dtype=np.float32) -> str:
        self.flip += 1
        self.inits.append(tensor_proto(name, np.ascontiguousarray(arr, dtype=dtype), raw=self.flip % 3 != 0))  # both encodings occur
        return name

    def node(self, op: str, inputs: List[str], attrs: Dict[str, object] = None) -> str:
        self.counter += 1
        out_name = f"/{op}_{self.counter}_output_0"
        body = b"".join(_ld(1, i.encode()) for i in inputs) + _ld(2, out_name.encode()) + _ld(3, f"/{op}_{self.counter}".encode())
        body += _ld(4, op.encode())
        for k, v in (attrs or {}).items():
            a = _ld(1, k.encode())
            if isinstance(v, float):
                a += _key(2, 5) + struct.pack("<f", v) + _key(20, 0) + _varint(1)
            elif isinstance(v, int):
                a += _key(3, 0) + _varint(v) + _key(20, 0) + _varint(2)
            elif isinstance(v, np.ndarray):
                a += _ld(5, tensor_proto("", v)) + _key(20, 0) + _varint(4)
            else:
                a += _ld(8, b"".join(_varint(int(x)) for x in v)) + _key(20, 0) + _varint(7)
            body += _ld(5, a)
        self.nodes.append(body)
        return out_name

    def scalar(self, value: float) -> str:
        return self.node("Constant", [], {"value": np.array(value, np.float32)})

    def model(self, inputs: List[str], outputs: List[str]) -> bytes:
        g = b"".join(_ld(1, n) for n in self.nodes) + _ld(2, b"main_graph") + b"".join(_ld(5, t) for t in self.inits)
        g += b"".join(_ld(11, _ld(1, i.encode())) for i in inputs) + b"".join(_ld(12, _ld(1, o.encode())) for o in outputs)
        opset = _ld(1, b"") + _key(2, 0) + _varint(17)
        return _key(1, 0) + _varint(8) + _ld(2, b"pytorch") + _ld(3, b"2.1.0") + _ld(7, g) + _ld(8, opset)

    # -- building blocks -----------------------------------------------------------------------------------
    def gelu(self, x: str) -> str:
        e = self.node("Erf", [self.node("Div", [x, self.scalar(1.4142135381698608)])])
        return self.node("Mul", [self.node("Mul", [x, self.node("Add", [e, self.scalar(1.0)])]), self.scalar(0.5)])

    def conv_bn(self, x: str, sd, prefix: str, groups: int = 1, stride: int = 1) -> str:
        w = sd[prefix + ".c.weight"]
        scale = sd[prefix + ".bn.weight"] / np.sqrt(sd[prefix + ".bn.running_var"] + BN_EPS)
        wf = w * scale[:, None, None, None]
        bf = sd[prefix + ".bn.bias"] - sd[prefix + ".bn.running_mean"] * scale
        k = w.shape[2]
        return self.node("Conv", [x, self.init(self.anon("Conv"), wf), self.init(self.anon("Conv"), bf)],
                         {"group": groups, "kernel_shape": [k, k], "pads": [k // 2] * 4, "strides": [stride, stride], "dilations": [1, 1]})

    def linear(self, x: str, sd, prefix: str, gemm: bool = False) -> str:
        w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
        if gemm:
            return self.node("Gemm", [x, self.init(prefix + ".weight", w), self.init(prefix + ".bias", b)], {"alpha": 1.0, "beta": 1.0, "transB": 1})
        y = self.node("MatMul", [x, self.init(self.anon("MatMul"), w.T)])
        return self.node("Add", [self.init(prefix + ".bias", b), y])

    def layernorm(self, x: str, sd, prefix: str) -> str:
        return self.node("LayerNormalization", [x, self.init(prefix + ".weight", sd[prefix + ".weight"]), self.init(prefix + ".bias", sd[prefix + ".bias"])],
                         {"axis": -1, "epsilon": 1e-5})

    def layernorm2d(self, x: str, sd, prefix: str) -> str:
        u = self.node("ReduceMean", [x], {"axes": [1], "keepdims": 1})
        d = self.node("Sub", [x, u])
        s = self.node("ReduceMean", [self.node("Pow", [d, self.scalar(2.0)])], {"axes": [1], "keepdims": 1})
        n = self.node("Div", [d, self.node("Sqrt", [self.node("Add", [s, self.scalar(1e-6)])])])
        c = sd[prefix + ".weight"].shape[0]
        y = self.node("Mul", [self.init(self.anon("Mul"), sd[prefix + ".weight"].reshape(c, 1, 1)), n])
        return self.node("Add", [y, self.init(self.anon("Add"), sd[prefix + ".bias"].reshape(c, 1, 1))])

    def reshape(self, x: str, shape) -> str:
        return self.node("Reshape", [x, self.init(self.anon("Reshape"), np.array(shape, np.int64), dtype=np.int64)])


def encoder_onnx(sd: Dict[str, np.ndarray]) -> bytes:
    e = Emitter()
    E = "image_encoder."
    x = e.node("Sub", ["input_image", e.init("pixel_mean", np.array([123.675, 116.28, 103.53]))])
    x = e.node("Div", [x, e.init("pixel_std", np.array([58.395, 57.12, 57.375]))])
    x = e.node("Transpose", [x], {"perm": [2, 0, 1]})
    x = e.node("Pad", [x, e.init(e.anon("Pad"), np.zeros(6, np.int64), dtype=np.int64)])
    x = e.gelu(e.conv_bn(x, sd, E + "patch_embed.seq.0", stride=2))
    x = e.conv_bn(x, sd, E + "patch_embed.seq.2", stride=2)
    for i in range(2):
        p = f"{E}layers.0.blocks.{i}"
        y = e.gelu(e.conv_bn(x, sd, p + ".conv1"))
        y = e.gelu(e.conv_bn(y, sd, p + ".conv2", groups=256))
        x = e.gelu(e.node("Add", [e.conv_bn(y, sd, p + ".conv3"), x]))

    def merge(x, i):
        p = f"{E}layers.{i}.downsample"
        y = e.gelu(e.conv_bn(x, sd, p + ".conv1"))
        y = e.gelu(e.conv_bn(y, sd, p + ".conv2", groups=DIMS[i + 1], stride=1 if DIMS[i + 1] == 320 else 2))
        return e.reshape(e.conv_bn(y, sd, p + ".conv3"), [1, DIMS[i + 1], -1])

    x = merge(x, 0)
    blk = 0
    for st in range(1, 4):
        C, heads, ws = DIMS[st], HEADS[st], WINDOWS[st]
        n = ws * ws
        idx, _ = attention_bias_idxs(ws)
        for i in range(DEPTHS[st]):
            p = f"{E}layers.{st}.blocks.{i}"
            y = e.node("Pad", [e.reshape(x, [1, -1, C]), e.init(e.anon("Pad"), np.zeros(8, np.int64), dtype=np.int64)])
            y = e.linear(e.layernorm(y, sd, p + ".attn.norm"), sd, p + ".attn.qkv")
            q = e.node("Mul", [e.reshape(y, [-1, n, heads, 96]), e.scalar(32 ** -0.5)])
            attn = e.node("MatMul", [q, e.node("Transpose", [y], {"perm": [0, 2, 3, 1]})])
            table = sd[p + ".attn.attention_biases"]
            if blk % 2 == 0:  # constant-folded gather: dense bias added directly
                attn = e.node("Add", [attn, e.init(e.anon("Add"), table[:, idx][None])])
            else:             # gather kept in the graph
                gathered = e.node("Gather", [e.init(p + ".attn.attention_biases", table), e.init(e.anon("Gather"), idx, dtype=np.int64)], {"axis": 1})
                attn = e.node("Add", [attn, gathered])
            blk += 1
            attn = e.node("MatMul", [e.node("Softmax", [attn], {"axis": -1}), y])
            x = e.node("Add", [x, e.linear(e.reshape(attn, [-1, n, C]), sd, p + ".attn.proj")])
            x = e.conv_bn(e.reshape(x, [1, C, 64, 64]), sd, p + ".local_conv", groups=C)
            m = e.gelu(e.linear(e.layernorm(x, sd, p + ".mlp.norm"), sd, p + ".mlp.fc1"))
            x = e.node("Add", [x, e.linear(m, sd, p + ".mlp.fc2")])
        if st < 3:
            x = merge(x, st)
    x = e.node("Conv", [e.reshape(x, [1, 320, 64, 64]), e.init(E + "neck.0.weight", sd[E + "neck.0.weight"])], {"kernel_shape": [1, 1]})
    x = e.layernorm2d(x, sd, E + "neck.1")
    x = e.node("Conv", [x, e.init(E + "neck.2.weight", sd[E + "neck.2.weight"])], {"kernel_shape": [3, 3], "pads": [1, 1, 1, 1]})
    x = e.layernorm2d(x, sd, E + "neck.3")
    return e.model(["input_image"], [x])


def decoder_onnx(sd: Dict[str, np.ndarray], dense_pe: np.ndarray) -> bytes:
    e = Emitter()
    P, D = "prompt_encoder.", "mask_decoder."
    c = e.node("Div", [e.node("Add", ["point_coords", e.scalar(0.5)]), e.scalar(1024.0)])
    c = e.node("MatMul", [e.node("Sub", [e.node("Mul", [c, e.scalar(2.0)]), e.scalar(1.0)]),
                          e.init(P + "pe_layer.positional_encoding_gaussian_matrix", sd[P + "pe_layer.positional_encoding_gaussian_matrix"])])
    pe = e.node("Concat", [e.node("Sin", [c]), e.node("Cos", [c])], {"axis": -1})
    pe = e.node("Add", [pe, e.node("Mul", [e.init(P + "not_a_point_embed.weight", sd[P + "not_a_point_embed.weight"]), e.node("Equal", ["point_labels", e.scalar(-1.0)])])])
    for i in range(4):
        pe = e.node("Add", [pe, e.node("Mul", [e.init(f"{P}point_embeddings.{i}.weight", sd[f"{P}point_embeddings.{i}.weight"]),
                                               e.node("Equal", ["point_labels", e.scalar(float(i))])])])
    m = "mask_input"
    for idx, nxt in ((0, 1), (3, 4)):
        w = sd[f"{P}mask_downscaling.{idx}.weight"]
        m = e.node("Conv", [m, e.init(f"{P}mask_downscaling.{idx}.weight", w), e.init(f"{P}mask_downscaling.{idx}.bias", sd[f"{P}mask_downscaling.{idx}.bias"])],
                   {"kernel_shape": [2, 2], "strides": [2, 2]})
        m = e.gelu(e.layernorm2d(m, sd, f"{P}mask_downscaling.{nxt}"))
    m = e.node("Conv", [m, e.init(P + "mask_downscaling.6.weight", sd[P + "mask_downscaling.6.weight"]),
                        e.init(P + "mask_downscaling.6.bias", sd[P + "mask_downscaling.6.bias"])], {"kernel_shape": [1, 1]})
    dense = e.node("Add", [e.node("Mul", ["has_mask_input", m]),
                           e.node("Mul", [e.node("Sub", [e.scalar(1.0), "has_mask_input"]), e.init(e.anon("Mul"), sd[P + "no_mask_embed.weight"].reshape(1, 256, 1, 1))])])
    toks = e.node("Expand", [e.init(e.anon("Expand"), np.concatenate([sd[D + "iou_token.weight"], sd[D + "mask_tokens.weight"]], 0)[None]),
                             e.init(e.anon("Expand"), np.array([1, 5, 256], np.int64), dtype=np.int64)])
    tokens = e.node("Concat", [toks, pe], {"axis": 1})
    src = e.node("Add", ["image_embeddings", dense])
    pos = e.init(e.anon("Add"), dense_pe.reshape(1, 256, 4096).transpose(0, 2, 1))  # folded constant: get_dense_pe()
    keys = e.node("Transpose", [e.reshape(src, [1, 256, 4096])], {"perm": [0, 2, 1]})

    def attn(q, k, v, prefix):
        qp = e.linear(q, sd, prefix + ".q_proj")
        kp = e.linear(k, sd, prefix + ".k_proj")
        vp = e.linear(v, sd, prefix + ".v_proj")
        a = e.node("Softmax", [e.node("Div", [e.node("MatMul", [qp, kp]), e.scalar(4.0)])], {"axis": -1})
        return e.linear(e.node("MatMul", [a, vp]), sd, prefix + ".out_proj")

    queries = tokens
    for i in range(2):
        p = f"{D}transformer.layers.{i}"
        q = queries if i == 0 else e.node("Add", [queries, tokens])
        a = attn(q, q, queries, p + ".self_attn")
        queries = e.layernorm(a if i == 0 else e.node("Add", [queries, a]), sd, p + ".norm1")
        q = e.node("Add", [queries, tokens])
        k = e.node("Add", [keys, pos])
        queries = e.layernorm(e.node("Add", [queries, attn(q, k, keys, p + ".cross_attn_token_to_image")]), sd, p + ".norm2")
        h = e.node("Relu", [e.linear(queries, sd, p + ".mlp.lin1")])
        queries = e.layernorm(e.node("Add", [queries, e.linear(h, sd, p + ".mlp.lin2")]), sd, p + ".norm3")
        q = e.node("Add", [queries, tokens])
        k = e.node("Add", [keys, pos])
        keys = e.layernorm(e.node("Add", [keys, attn(k, q, queries, p + ".cross_attn_image_to_token")]), sd, p + ".norm4")
    q = e.node("Add", [queries, tokens])
    k = e.node("Add", [keys, pos])
    queries = e.layernorm(e.node("Add", [queries, attn(q, k, keys, D + "transformer.final_attn_token_to_image")]), sd, D + "transformer.norm_final_attn")
    up = e.node("ConvTranspose", [e.reshape(keys, [1, 256, 64, 64]), e.init(D + "output_upscaling.0.weight", sd[D + "output_upscaling.0.weight"]),
                                  e.init(D + "output_upscaling.0.bias", sd[D + "output_upscaling.0.bias"])], {"kernel_shape": [2, 2], "strides": [2, 2]})
    up = e.gelu(e.layernorm2d(up, sd, D + "output_upscaling.1"))
    up = e.gelu(e.node("ConvTranspose", [up, e.init(D + "output_upscaling.3.weight", sd[D + "output_upscaling.3.weight"]),
                                         e.init(D + "output_upscaling.3.bias", sd[D + "output_upscaling.3.bias"])], {"kernel_shape": [2, 2], "strides": [2, 2]}))
    hyper = []
    for mi in range(4):
        t = e.node("Gather", [queries, e.init(e.anon("Gather"), np.array(1 + mi, np.int64), dtype=np.int64)], {"axis": 1})
        for j in range(3):
            t = e.linear(t, sd, f"{D}output_hypernetworks_mlps.{mi}.layers.{j}", gemm=True)
            if j < 2:
                t = e.node("Relu", [t])
        hyper.append(t)
    masks = e.node("MatMul", [e.node("Concat", hyper, {"axis": 1}), e.reshape(up, [1, 32, 65536])])
    t = e.node("Gather", [queries, e.init(e.anon("Gather"), np.array(0, np.int64), dtype=np.int64)], {"axis": 1})
    for j in range(3):
        t = e.linear(t, sd, f"{D}iou_prediction_head.layers.{j}", gemm=True)
        if j < 2:
            t = e.node("Relu", [t])
    masks = e.node("Resize", [masks, e.init(e.anon("Resize"), np.array([1.0, 1.0, 4.0, 4.0], np.float32))], {"mode": 1})
    return e.model(["image_embeddings", "point_coords", "point_labels", "mask_input", "has_mask_input", "orig_im_size"], [masks, t])
