"""Reads the weights of the reference's three MobileSAM ONNX graphs back into MobileSAM state-dict names.

The reference ships `mobile_sam_image_encoder.onnx`, `sam_mask_decoder_single.onnx` and `sam_mask_decoder_multi.onnx`
(models/segmentation/CMakeLists.txt:2-16, exported by script/export_models.py:21-43 with opset 17).  Neither `onnx` nor
`onnxruntime` exists here, so this module carries its own reader for the protobuf wire format (just the handful of ONNX
messages that hold tensors) and recovers the parameters by WALKING THE GRAPH IN EXECUTION ORDER: torch's exporter folds
BatchNorm into the convolutions, transposes Linear weights and gives both anonymous names (`onnx::Conv_1234`,
`onnx::MatMul_987`), so names cannot be trusted -- but the order in which weight-carrying nodes execute, and the shape of
every weight, are fixed by the architecture.  Every step checks the operator kind and the shape and
fails loudly on the first mismatch.

Folded Conv+BN pairs are stored as conv weight + identity BatchNorm (gamma 1, running_var 1 - eps, running_mean 0, beta =
the folded bias), so the engine's and the oracle's loaders, which fold BatchNorm themselves, reproduce the graph's numbers.

PARITY UNPINNED: no real .onnx file is available offline; the importer is exercised on graphs written by
tests/onnx_emit.py from the oracle's state dict, which follow the exporter conventions described above.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

ENCODER_ONNX = "mobile_sam_image_encoder.onnx"
DECODER_ONNX = ("sam_mask_decoder_single.onnx", "sam_mask_decoder_multi.onnx")
MD5 = {ENCODER_ONNX: "9E0ED7F27DC33C6DFD08A0CBA6EAC141", "sam_mask_decoder_multi.onnx": "CFF1C936628337B5F4D4EFAD9F94CCA7",
       "sam_mask_decoder_single.onnx": "5A5174CCF1A62EC4FFF38E2ACBBD8201"}  # models/segmentation/CMakeLists.txt:5,10,15

DIMS = (64, 128, 160, 320)
DEPTHS = (2, 2, 6, 2)
HEADS = (2, 4, 5, 10)
WINDOWS = (7, 7, 14, 7)
BN_EPS = 1e-5


# ---- protobuf wire format ------------------------------------------------------------------------
def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf: memoryview):
    """Yields (field number, wire type, value) of one message; length-delimited values come as memoryviews."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = bytes(buf[pos:pos + 8]); pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4]); pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        if pos > n:
            raise ValueError("truncated protobuf message")
        yield field, wt, v


def _packed_varints(v, wt) -> List[int]:
    if wt == 0:
        return [v]
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(x)
    return out


_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64, 9: np.bool_, 10: np.float16, 11: np.float64}


def _tensor(buf: memoryview) -> Tuple[str, np.ndarray]:
    """TensorProto: dims=1, data_type=2, float_data=4, int32_data=5, int64_data=7, name=8, raw_data=9, double_data=10."""
    dims: List[int] = []
    dtype = 1
    name = ""
    raw = None
    floats: List[float] = []
    ints: List[int] = []
    for f, wt, v in _fields(buf):
        if f == 1:
            dims += _packed_varints(v, wt)
        elif f == 2:
            dtype = v
        elif f == 8:
            name = bytes(v).decode()
        elif f == 9:
            raw = bytes(v)
        elif f == 4:
            floats += list(np.frombuffer(bytes(v), "<f4")) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif f in (5, 7):
            ints += _packed_varints(v, wt)
        elif f == 10:
            floats += list(np.frombuffer(bytes(v), "<f8")) if wt == 2 else [struct.unpack("<d", v)[0]]
        elif f == 13 or (f == 14 and v == 1):
            raise ValueError(f"tensor {name}: external data is not supported")
    if dtype not in _DTYPES:
        raise ValueError(f"tensor {name}: unsupported data type {dtype}")
    np_t = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, np.dtype(np_t).newbyteorder("<")).astype(np_t)
    elif floats:
        arr = np.array(floats, np_t)
    else:
        ints = [x - (1 << 64) if x >= (1 << 63) else x for x in ints]  # negative int64 travel as 64-bit two's complement
        arr = np.array(ints, np_t)
    return name, arr.reshape(dims) if dims else arr.reshape(())


class Node:
    def __init__(self):
        self.op = ""
        self.name = ""
        self.inputs: List[str] = []
        self.outputs: List[str] = []
        self.attrs: Dict[str, object] = {}


def _node(buf: memoryview) -> Node:
    """NodeProto: input=1, output=2, name=3, op_type=4, attribute=5.  AttributeProto: name=1, f=2, i=3, t=5, ints=8."""
    n = Node()
    for f, wt, v in _fields(buf):
        if f == 1:
            n.inputs.append(bytes(v).decode())
        elif f == 2:
            n.outputs.append(bytes(v).decode())
        elif f == 3:
            n.name = bytes(v).decode()
        elif f == 4:
            n.op = bytes(v).decode()
        elif f == 5:
            an, val = "", None
            for af, awt, av in _fields(v):
                if af == 1:
                    an = bytes(av).decode()
                elif af == 2:
                    val = struct.unpack("<f", av)[0]
                elif af == 3:
                    val = av - (1 << 64) if av >= (1 << 63) else av
                elif af == 5:
                    val = _tensor(av)[1]
                elif af == 8:
                    val = (val or []) + _packed_varints(av, awt)
            n.attrs[an] = val
    return n


class Graph:
    def __init__(self):
        self.nodes: List[Node] = []
        self.initializers: Dict[str, np.ndarray] = {}
        self.inputs: List[str] = []
        self.outputs: List[str] = []


def parse_model(data: bytes) -> Graph:
    """ModelProto.graph = 7; GraphProto: node=1, initializer=5, input=11, output=12."""
    g = Graph()
    graph_buf = None
    for f, wt, v in _fields(memoryview(data)):
        if f == 7 and wt == 2:
            graph_buf = v
    if graph_buf is None:
        raise ValueError("not an ONNX model: no graph")
    for f, wt, v in _fields(graph_buf):
        if f == 1:
            g.nodes.append(_node(v))
        elif f == 5:
            name, arr = _tensor(v)
            g.initializers[name] = arr
        elif f in (11, 12):
            for vf, _, vv in _fields(v):
                if vf == 1:
                    (g.inputs if f == 11 else g.outputs).append(bytes(vv).decode())
    for n in g.nodes:  # Constant nodes are initializers under another name
        if n.op == "Constant" and isinstance(n.attrs.get("value"), np.ndarray):
            g.initializers[n.outputs[0]] = n.attrs["value"]
    return g


def load_graph(path: str) -> Graph:
    with open(path, "rb") as f:
        return parse_model(f.read())


# ---- weight-carrying nodes in execution order -------------------------------------------------------
class Event:
    def __init__(self, op: str, tensors: List[np.ndarray], node: Node):
        self.op, self.tensors, self.node = op, tensors, node

    def __repr__(self):
        return f"{self.op}{[tuple(t.shape) for t in self.tensors]}"


def weight_events(g: Graph) -> List[Event]:
    """Nodes that consume a floating-point initializer with more than one element, in graph (= execution) order."""
    out = []
    for n in g.nodes:
        if n.op == "Constant":
            continue
        ts = [g.initializers[i] for i in n.inputs if i in g.initializers]
        ts = [t for t in ts if t.dtype in (np.float32, np.float16, np.float64) and t.size > 1]
        if ts:
            out.append(Event(n.op, [t.astype(np.float32) for t in ts], n))
    return out


class _Cursor:
    def __init__(self, events: List[Event], what: str):
        self.ev, self.i, self.what = events, 0, what

    def fail(self, msg):
        ctx = ", ".join(repr(e) for e in self.ev[max(0, self.i - 2): self.i + 3])
        raise ValueError(f"{self.what}: {msg} at weight-carrying node {self.i} (around: {ctx})")

    def peek(self) -> Optional[Event]:
        return self.ev[self.i] if self.i < len(self.ev) else None

    def next(self, ops, why) -> Event:
        e = self.peek()
        if e is None or e.op not in ops:
            self.fail(f"expected {'/'.join(ops)} for {why}, found {e}")
        self.i += 1
        return e

    def skip_until(self, op):
        while self.peek() is not None and self.peek().op != op:
            self.i += 1

    # -- typed readers ---------------------------------------------------------------------------------
    def conv(self, out: Dict[str, np.ndarray], prefix: str, shape, bias=True, bn=True):
        e = self.next(("Conv",), prefix)
        w = e.tensors[0]
        if tuple(w.shape) != tuple(shape):
            self.fail(f"{prefix}: conv weight {w.shape}, expected {shape}")
        b = e.tensors[1] if len(e.tensors) > 1 else None
        if bias and (b is None or b.shape != (shape[0],)):
            self.fail(f"{prefix}: conv bias missing")
        if bn:  # Conv2d_BN folded by the exporter -> conv + identity BatchNorm carrying the folded bias
            out[prefix + ".c.weight"] = w
            out[prefix + ".bn.weight"] = np.ones(shape[0], np.float32)
            out[prefix + ".bn.bias"] = b if b is not None else np.zeros(shape[0], np.float32)
            out[prefix + ".bn.running_mean"] = np.zeros(shape[0], np.float32)
            out[prefix + ".bn.running_var"] = np.full(shape[0], 1.0 - BN_EPS, np.float32)
        else:
            out[prefix + ".weight"] = w
            if b is not None:
                out[prefix + ".bias"] = b

    def linear(self, out, prefix: str, n: int, k: int):
        """nn.Linear(k, n): Gemm(W (n,k) transB=1, b) or MatMul(W^T (k,n)) followed by Add(b)."""
        e = self.next(("Gemm", "MatMul"), prefix)
        w = e.tensors[0]
        if e.op == "Gemm":
            w = w if int(e.node.attrs.get("transB", 0) or 0) else w.T
            if len(e.tensors) < 2:
                self.fail(f"{prefix}: Gemm without bias")
            b = e.tensors[1]
        else:
            w = w.T
            b = self.next(("Add",), prefix + ".bias").tensors[0]
        if w.shape != (n, k) or b.reshape(-1).shape != (n,):
            self.fail(f"{prefix}: linear weight {w.shape} / bias {b.shape}, expected {(n, k)}")
        out[prefix + ".weight"] = np.ascontiguousarray(w)
        out[prefix + ".bias"] = b.reshape(-1)

    def layernorm(self, out, prefix: str, c: int):
        """LayerNormalization(scale, bias) (opset 17), or the decomposed form ... Mul(weight) Add(bias) (LayerNorm2d)."""
        e = self.next(("LayerNormalization", "Mul"), prefix)
        if e.op == "LayerNormalization":
            if len(e.tensors) != 2:
                self.fail(f"{prefix}: LayerNormalization without scale and bias")
            w, b = e.tensors
        else:
            w = e.tensors[0]
            b = self.next(("Add",), prefix + ".bias").tensors[0]
        if w.size != c or b.size != c:
            self.fail(f"{prefix}: norm of {w.size} channels, expected {c}")
        out[prefix + ".weight"] = w.reshape(-1)
        out[prefix + ".bias"] = b.reshape(-1)


def attention_bias_idxs(ws: int) -> Tuple[np.ndarray, int]:
    """TinyViT Attention.attention_bias_idxs: offsets (|dy|, |dx|) in first-seen order."""
    n = ws * ws
    offsets: Dict[Tuple[int, int], int] = {}
    idx = np.zeros((n, n), np.int64)
    for a in range(n):
        for b in range(n):
            o = (abs(a // ws - b // ws), abs(a % ws - b % ws))
            if o not in offsets:
                offsets[o] = len(offsets)
            idx[a, b] = offsets[o]
    return idx, len(offsets)


def encoder_state(g: Graph) -> Dict[str, np.ndarray]:
    """mobile_sam_image_encoder.onnx -> `image_encoder.*` tensors."""
    c = _Cursor(weight_events(g), ENCODER_ONNX)
    out: Dict[str, np.ndarray] = {}
    E = "image_encoder."
    c.skip_until("Conv")  # the preprocessing in front (mean / std constants) is fixed by the engine
    c.conv(out, E + "patch_embed.seq.0", (32, 3, 3, 3))
    c.conv(out, E + "patch_embed.seq.2", (64, 32, 3, 3))
    for i in range(2):
        p = f"{E}layers.0.blocks.{i}"
        c.conv(out, p + ".conv1", (256, 64, 1, 1))
        c.conv(out, p + ".conv2", (256, 1, 3, 3))
        c.conv(out, p + ".conv3", (64, 256, 1, 1))

    def merge(i):
        p = f"{E}layers.{i}.downsample"
        c.conv(out, p + ".conv1", (DIMS[i + 1], DIMS[i], 1, 1))
        c.conv(out, p + ".conv2", (DIMS[i + 1], 1, 3, 3))
        c.conv(out, p + ".conv3", (DIMS[i + 1], DIMS[i + 1], 1, 1))

    merge(0)
    for st in range(1, 4):
        C, heads, ws = DIMS[st], HEADS[st], WINDOWS[st]
        n = ws * ws
        idx, n_off = attention_bias_idxs(ws)
        for i in range(DEPTHS[st]):
            p = f"{E}layers.{st}.blocks.{i}"
            c.layernorm(out, p + ".attn.norm", C)
            c.linear(out, p + ".attn.qkv", 3 * C, C)
            # relative-position bias: Gather(attention_biases, idxs) or, constant-folded, Add(dense (1?, heads, n, n))
            e = c.next(("Gather", "Add"), p + ".attn.attention_biases")
            t = e.tensors[0]
            if e.op == "Gather" and t.shape == (heads, n_off):
                table = t
            elif t.size == heads * n * n:
                dense = t.reshape(heads, n, n)
                table = np.zeros((heads, n_off), np.float32)
                seen = np.zeros(n_off, bool)
                for a in range(n):
                    for b in range(n):
                        o = idx[a, b]
                        if not seen[o]:
                            table[:, o] = dense[:, a, b]
                            seen[o] = True
                if not np.array_equal(table[:, idx], dense):
                    c.fail(f"{p}: dense attention bias is not a gather of a (heads, offsets) table")
            else:
                c.fail(f"{p}: attention bias of shape {t.shape}")
            out[p + ".attn.attention_biases"] = table
            c.linear(out, p + ".attn.proj", C, C)
            c.conv(out, p + ".local_conv", (C, 1, 3, 3))
            c.layernorm(out, p + ".mlp.norm", C)
            c.linear(out, p + ".mlp.fc1", 4 * C, C)
            c.linear(out, p + ".mlp.fc2", C, 4 * C)
        if st < 3:
            merge(st)
    c.conv(out, E + "neck.0", (256, 320, 1, 1), bias=False, bn=False)
    c.layernorm(out, E + "neck.1", 256)
    c.conv(out, E + "neck.2", (256, 256, 3, 3), bias=False, bn=False)
    c.layernorm(out, E + "neck.3", 256)
    if c.peek() is not None:
        c.fail("unexpected weights after the neck")
    return out


def decoder_state(g: Graph, name: str = DECODER_ONNX[1]) -> Dict[str, np.ndarray]:
    """sam_mask_decoder_{single,multi}.onnx -> `prompt_encoder.*` and `mask_decoder.*` tensors (both files hold the same)."""
    c = _Cursor(weight_events(g), name)
    out: Dict[str, np.ndarray] = {}
    P, D = "prompt_encoder.", "mask_decoder."
    # SamOnnxModel._embed_points: coords @ gaussian matrix, then the label embeddings
    e = c.next(("MatMul",), "positional_encoding_gaussian_matrix")
    if e.tensors[0].shape != (2, 128):
        c.fail(f"gaussian matrix of shape {e.tensors[0].shape}")
    out[P + "pe_layer.positional_encoding_gaussian_matrix"] = e.tensors[0]
    out[P + "not_a_point_embed.weight"] = c.next(("Mul",), "not_a_point_embed").tensors[0].reshape(1, 256)
    for i in range(4):
        out[f"{P}point_embeddings.{i}.weight"] = c.next(("Mul",), f"point_embeddings.{i}").tensors[0].reshape(1, 256)
    # SamOnnxModel._embed_masks: mask_downscaling (never influences the result with the reference's inputs), no_mask_embed
    c.conv(out, P + "mask_downscaling.0", (4, 1, 2, 2), bn=False)
    c.layernorm(out, P + "mask_downscaling.1", 4)
    c.conv(out, P + "mask_downscaling.3", (16, 4, 2, 2), bn=False)
    c.layernorm(out, P + "mask_downscaling.4", 16)
    c.conv(out, P + "mask_downscaling.6", (256, 16, 1, 1), bn=False)
    out[P + "no_mask_embed.weight"] = c.next(("Mul",), "no_mask_embed").tensors[0].reshape(1, 256)
    # MaskDecoder.predict_masks: [iou_token; mask_tokens] concatenated (constant-folded to one (5, 256) initializer or kept
    # as two), consumed by the Expand / Concat in front of the transformer; the dense PE (1, 256, 64, 64) is a folded constant
    toks = []
    while sum(t.shape[0] for t in toks) < 5:
        e = c.next(("Concat", "Expand", "Unsqueeze", "Tile", "Add", "Reshape", "Gather"), "output tokens / dense positional encoding")
        for t in e.tensors:
            if t.ndim >= 2 and t.shape[-1] == 256 and t.size in (256, 1024, 1280):
                toks.append(t.reshape(-1, 256))
    tok = np.concatenate(toks, 0)
    out[D + "iou_token.weight"] = tok[:1]
    out[D + "mask_tokens.weight"] = tok[1:5]

    def skip_constants():  # folded dense positional encoding and its derivatives carry no parameter
        while c.peek() is not None and c.peek().op not in ("MatMul", "Gemm", "LayerNormalization", "ConvTranspose", "Conv"):
            c.i += 1

    def attn(prefix, dim, internal):
        for nme in ("q_proj", "k_proj", "v_proj"):
            skip_constants()
            c.linear(out, f"{prefix}.{nme}", internal, dim)
        skip_constants()
        c.linear(out, prefix + ".out_proj", dim, internal)

    for i in range(2):
        p = f"{D}transformer.layers.{i}"
        attn(p + ".self_attn", 256, 256)
        c.layernorm(out, p + ".norm1", 256)
        attn(p + ".cross_attn_token_to_image", 256, 128)
        c.layernorm(out, p + ".norm2", 256)
        c.linear(out, p + ".mlp.lin1", 2048, 256)
        c.linear(out, p + ".mlp.lin2", 256, 2048)
        c.layernorm(out, p + ".norm3", 256)
        attn(p + ".cross_attn_image_to_token", 256, 128)
        c.layernorm(out, p + ".norm4", 256)
    attn(D + "transformer.final_attn_token_to_image", 256, 128)
    c.layernorm(out, D + "transformer.norm_final_attn", 256)
    e = c.next(("ConvTranspose",), "output_upscaling.0")
    if e.tensors[0].shape != (256, 64, 2, 2) or len(e.tensors) < 2:
        c.fail(f"output_upscaling.0 of shape {e.tensors[0].shape}")
    out[D + "output_upscaling.0.weight"], out[D + "output_upscaling.0.bias"] = e.tensors[0], e.tensors[1]
    c.layernorm(out, D + "output_upscaling.1", 64)
    e = c.next(("ConvTranspose",), "output_upscaling.3")
    if e.tensors[0].shape != (64, 32, 2, 2) or len(e.tensors) < 2:
        c.fail(f"output_upscaling.3 of shape {e.tensors[0].shape}")
    out[D + "output_upscaling.3.weight"], out[D + "output_upscaling.3.bias"] = e.tensors[0], e.tensors[1]
    for m in range(4):
        for j, (n, k) in enumerate(((256, 256), (256, 256), (32, 256))):
            c.linear(out, f"{D}output_hypernetworks_mlps.{m}.layers.{j}", n, k)
    for j, (n, k) in enumerate(((256, 256), (256, 256), (4, 256))):
        c.linear(out, f"{D}iou_prediction_head.layers.{j}", n, k)
    return out


def state_from_onnx_dir(directory: str) -> Dict[str, np.ndarray]:
    """<directory>/{mobile_sam_image_encoder, sam_mask_decoder_multi | _single}.onnx -> one MobileSAM state dict."""
    import os
    state = encoder_state(load_graph(os.path.join(directory, ENCODER_ONNX)))
    for name in (DECODER_ONNX[1], DECODER_ONNX[0]):
        path = os.path.join(directory, name)
        if os.path.exists(path):
            state.update(decoder_state(load_graph(path), name))
            return state
    raise FileNotFoundError(f"no {DECODER_ONNX[0]} / {DECODER_ONNX[1]} in {directory}")
