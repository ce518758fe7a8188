"""Writer/reader for the engine's flat weight container ("DLIMGB2", see csrc/weights.hpp).

The container keeps the MobileSAM state-dict names, so converting a real checkpoint is a 1:1 dump
(tools/convert_checkpoint.py).  Pure numpy; no torch required.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

MAGIC = b"DLIMGB2\0"
WEIGHT_FILE_NAME = "mobile_sam_b200.bin"  # lives in <model_directory>/segmentation/


def save(path: str, tensors: Dict[str, np.ndarray]) -> None:
    names = sorted(tensors)
    header = bytearray()
    header += MAGIC
    header += struct.pack("<II", 1, len(names))
    offset = 0
    blobs = []
    for name in names:
        a = np.ascontiguousarray(tensors[name], dtype=np.float32)
        nb = name.encode()
        header += struct.pack("<H", len(nb)) + nb
        header += struct.pack("<B", a.ndim)
        header += struct.pack("<" + "I" * a.ndim, *a.shape)
        header += struct.pack("<QQ", offset, a.size)
        blobs.append(a.tobytes())
        offset += a.size * 4
    with open(path, "wb") as f:
        f.write(bytes(header))
        for b in blobs:
            f.write(b)


def load(path: str) -> Dict[str, np.ndarray]:
    with open(path, "rb") as f:
        data = f.read()
    assert data[:8] == MAGIC, "not a DLIMGB2 container"
    version, count = struct.unpack_from("<II", data, 8)
    assert version == 1
    pos = 16
    recs = []
    for _ in range(count):
        (ln,) = struct.unpack_from("<H", data, pos); pos += 2
        name = data[pos:pos + ln].decode(); pos += ln
        (nd,) = struct.unpack_from("<B", data, pos); pos += 1
        shape = struct.unpack_from("<" + "I" * nd, data, pos); pos += 4 * nd
        off, numel = struct.unpack_from("<QQ", data, pos); pos += 16
        recs.append((name, shape, off, numel))
    out = {}
    for name, shape, off, numel in recs:
        out[name] = np.frombuffer(data, np.float32, numel, pos + off).reshape(shape).copy()
    return out


def from_state_dict(sd) -> Dict[str, np.ndarray]:
    """torch state_dict -> {name: float32 ndarray}; integer buffers (num_batches_tracked, bias idxs) are dropped."""
    out = {}
    for k, v in sd.items():
        if not v.dtype.is_floating_point:
            continue
        out[k] = v.detach().cpu().float().numpy()
    return out
