"""Development probe: where does the end-to-end step time go (uploads / encoder / downloads / host overhead)?"""
import os, sys, time, tempfile, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
B = int(os.environ.get("B", "16"))
os.environ.setdefault("DLIMG_B200_MAX_BATCH", str(B))
import dlimgedit_b200 as dl
from dlimgedit_b200 import synthetic_weights
d = tempfile.mkdtemp(); synthetic_weights.write_model_dir(d, seed=0)
env = dl.Environment(dl.Options(dl.Backend.gpu, d))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); env.set_stream(stream.cuda_stream)
rng = np.random.default_rng(0)
host = [torch.from_numpy(rng.integers(0, 256, (B, 1024, 1024, 4), dtype=np.uint8)).pin_memory() for _ in range(4)]
dev = [h.cuda() for h in host]
outs = [torch.empty(B, 256, 64, 64).pin_memory() for _ in range(3)]
ext = dl.Extent(1024, 1024)
keep = collections.deque(maxlen=2)

def run(name, fn, steps=12):
    for i in range(3): fn(i)
    env.synchronize(); torch.cuda.synchronize()
    t0 = time.perf_counter(); th = 0.0
    for i in range(steps):
        a = time.perf_counter(); fn(i); th += time.perf_counter() - a
    env.synchronize(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name:44s} {dt / steps * 1e3:7.3f} ms/step  host-side {th / steps * 1e3:7.3f} ms/step  {B * steps / dt:8.1f} img/s", flush=True)
    keep.clear()

def hv(i): return [dl.ImageView(host[i % 4][j].numpy(), ext, dl.Channels.rgba) for j in range(B)]
def dv(i): return [dl.ImageView(dev[i % 4][j].data_ptr(), ext, dl.Channels.rgba, device=True) for j in range(B)]
def f_dev(i): keep.append(env.process_batch(dv(i)))
def f_host(i): keep.append(env.process_batch(hv(i)))
def f_host_async(i):
    segs = env.process_batch(hv(i)); keep.append(segs)
    for j, s in enumerate(segs): s.embedding_async(outs[i % 3][j].numpy())
def f_dev_async(i):
    segs = env.process_batch(dv(i)); keep.append(segs)
    for j, s in enumerate(segs): s.embedding_async(outs[i % 3][j].numpy())
def f_host_sync(i):
    segs = env.process_batch(hv(i)); keep.append(segs)
    for j, s in enumerate(segs): s.embedding(out=outs[i % 3][j].numpy())
def f_views_only(i): hv(i)
run("device pixels, no read", f_dev)
run("host pixels, no read", f_host)
run("device pixels, async embedding read", f_dev_async)
run("host pixels, async embedding read", f_host_async)
run("host pixels, blocking embedding read", f_host_sync)
run("building host views only (python)", f_views_only)
