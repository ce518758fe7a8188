"""Development probe: the stand-alone pre/post stages of BASELINE config 4 (4K resize RGB / BGRA, mask upsample to 4K and to
1024^2 with 64 planes), a few launches each, for ncu."""
import ctypes, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dlimgedit_b200 as dl
from dlimgedit_b200 import synthetic_weights
d = tempfile.mkdtemp(); synthetic_weights.write_model_dir(d, seed=0)
env = dl.Environment(dl.Options(dl.Backend.gpu, d))
rng = np.random.default_rng(2)
ext = (ctypes.c_int * 2)()
for ch, bpp in ((dl.Channels.rgb, 3), (dl.Channels.bgra, 4)):
    img = torch.from_numpy(rng.integers(0, 256, (2160, 3840 * bpp), dtype=np.uint8)).cuda()
    out = torch.empty(1024 * 576 * bpp, dtype=torch.uint8, device="cuda")
    v = dl.ImageView(img.data_ptr(), dl.Extent(3840, 2160), ch, 3840 * bpp, device=True).to_c()
    for _ in range(3):
        assert dl.ext().resize_longest_side(env.handle(), ctypes.byref(v), 1024, out.data_ptr(), ext) == 0
for (w, h, cnt) in ((3840, 2160, 16), (1024, 1024, 64), (1800, 1200, 16)):
    low = torch.randn(cnt, 256, 256, device="cuda")
    o = torch.empty(cnt, h, w, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        assert dl.ext().mask_postprocess(env.handle(), low.data_ptr(), cnt, w, h, o.data_ptr()) == 0
env.synchronize()
print("ok")
