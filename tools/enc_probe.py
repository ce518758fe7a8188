"""Development probe: one 32-image encoder pass (eager, graphs off) for ncu."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
B = int(os.environ.get("B", "32"))
os.environ.setdefault("DLIMG_B200_MAX_BATCH", str(B))
os.environ.setdefault("DLIMG_B200_GRAPHS", "0")
import dlimgedit_b200 as dl
from dlimgedit_b200 import synthetic_weights
d = tempfile.mkdtemp(); synthetic_weights.write_model_dir(d, seed=0)
env = dl.Environment(dl.Options(dl.Backend.gpu, d))
rng = np.random.default_rng(0)
imgs = [torch.from_numpy(rng.integers(0, 256, (1024, 1024, 4), dtype=np.uint8)).cuda() for _ in range(B)]
views = [dl.ImageView(t.data_ptr(), dl.Extent(1024, 1024), dl.Channels.rgba, device=True) for t in imgs]
for _ in range(int(os.environ.get("CALLS", "2"))):
    segs = env.process_batch(views)
env.synchronize()
print("ok", len(segs))
