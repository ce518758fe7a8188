import os, sys, time, tempfile
sys.path.insert(0, '/root/repo')
import numpy as np, torch
os.environ.setdefault("DLIMG_B200_MAX_BATCH", "8")
import dlimgedit_b200 as dl
from dlimgedit_b200 import synthetic_weights
d = tempfile.mkdtemp(); synthetic_weights.write_model_dir(d, seed=0)
env = dl.Environment(dl.Options(dl.Backend.gpu, d))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); env.set_stream(stream.cuda_stream)
rng = np.random.default_rng(0)
sets = [torch.from_numpy(rng.integers(0, 256, (8, 1024, 1024, 4), dtype=np.uint8)).cuda() for _ in range(6)]
ext = dl.Extent(1024, 1024)
keep = []
for i in range(16):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    views = [dl.ImageView(sets[i % 6][j].data_ptr(), ext, dl.Channels.rgba, device=True) for j in range(8)]
    keep.append(env.process_batch(views))
    t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(i, 'host ms %.2f total ms %.2f' % ((t1 - t0) * 1e3, (t2 - t0) * 1e3), 'mem GB %.2f' % (torch.cuda.mem_get_info()[0] / 1e9))
    if i == 8: keep.clear()
