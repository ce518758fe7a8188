import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with `pytest -m gpu`)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_sam():
    """Seeded synthetic MobileSAM (oracle/mobile_sam_ref.py); the same weights are written for the engine."""
    import torch
    from oracle.mobile_sam_ref import build_synthetic
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    return build_synthetic(0)


@pytest.fixture(scope="session")
def model_dir(tmp_path_factory):
    """<dir>/segmentation/mobile_sam_b200.bin holding the seeded synthetic weights (the oracle loads the same)."""
    from dlimgedit_b200 import synthetic_weights
    d = tmp_path_factory.mktemp("models")
    synthetic_weights.write_model_dir(str(d), seed=0)
    return str(d)


@pytest.fixture(scope="session", autouse=True)
def _device_fills_complete_before_use():
    """The tests hand the library device buffers that torch has just filled (torch.zeros(..., device="cuda") for outputs) on
    torch's stream -- in a pytest process the legacy default stream, whose handle 0 the C ABI reads as "the environment's own
    stream".  That stream is non-blocking, so a library kernel could run before or UNDER the pending fill (seen once as ~2000
    cleared mask pixels in test_gpu_prepost).  Same contract as for any caller -- share a stream or synchronize -- and the
    tests synchronize: device fills return only when they have completed.  (Host-to-device copies of pageable tensors,
    `.cuda()`, are complete on return anyway; the debug-table kernels run on the legacy stream itself.)"""
    if not _has_gpu():
        yield
        return
    import torch
    originals = {name: getattr(torch, name) for name in ("zeros", "ones", "full", "zeros_like", "ones_like", "full_like")}

    def synchronous(fn):
        def wrapped(*args, **kwargs):
            t = fn(*args, **kwargs)
            if isinstance(t, torch.Tensor) and t.is_cuda:
                torch.cuda.synchronize()
            return t
        return wrapped

    for name, fn in originals.items():
        setattr(torch, name, synchronous(fn))
    yield
    for name, fn in originals.items():
        setattr(torch, name, fn)


@pytest.fixture(scope="session")
def env(model_dir):
    import dlimgedit_b200 as dl
    e = dl.Environment(dl.Options(dl.Backend.gpu, model_dir))
    # (torch's current stream in a pytest process is the legacy default stream, handle 0 = "the environment's own stream" for
    # the C ABI: the library works on its own non-blocking stream in these tests; see _device_fills_complete_before_use)
    import torch
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    yield e
    e.close()


def synthetic_image(h, w, c, seed):
    """Smooth-ish synthetic picture (blobs + noise) so that resampling and the network see structure."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, c), np.float32)
    for ch in range(c):
        acc = np.zeros((h, w), np.float32)
        for _ in range(6):
            cx, cy = rng.uniform(0, w), rng.uniform(0, h)
            s = rng.uniform(0.05, 0.3) * max(h, w)
            acc += rng.uniform(-1, 1) * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
        img[..., ch] = acc
    img = (img - img.min()) / (img.max() - img.min() + 1e-6) * 255
    img += rng.normal(0, 6, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)
