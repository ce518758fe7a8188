"""Development probe: one cached embedding, P point prompts per decoder call (masks stay on the device)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
P = int(os.environ.get("P", "64"))
os.environ.setdefault("DLIMG_B200_MAX_PROMPTS", str(P))
os.environ.setdefault("DLIMG_B200_MAX_BATCH", "1")
import dlimgedit_b200 as dl
from dlimgedit_b200 import synthetic_weights
d = tempfile.mkdtemp(); synthetic_weights.write_model_dir(d, seed=0)
env = dl.Environment(dl.Options(dl.Backend.gpu, d))
rng = np.random.default_rng(0)
img = torch.from_numpy(rng.integers(0, 256, (1024, 1024, 4), dtype=np.uint8)).cuda()
seg = env.process_batch([dl.ImageView(img.data_ptr(), dl.Extent(1024, 1024), dl.Channels.rgba, device=True)])[0]
prompts = [dl.Point(int(rng.integers(0, 1024)), int(rng.integers(0, 1024))) for _ in range(P)]
masks = torch.empty(P, 1024, 1024, dtype=torch.uint8, device="cuda")
ious = torch.empty(P, device="cuda")
ptrs = [masks[i].data_ptr() for i in range(P)]
for _ in range(int(os.environ.get("CALLS", "3"))):
    env.compute_masks_batch([seg] * P, prompts, multi=False, masks_out=ptrs, ious_out=ious.data_ptr())
env.synchronize()
print("ok", float(ious.mean()))
