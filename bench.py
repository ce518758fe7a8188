#!/usr/bin/env python
"""bench.py -- MobileSAM segmentation hot path on B200 (BASELINE.json configs 1-5 on one JSON line).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--prompts P] [--impl ours|reference] [--quick]

Headline (BASELINE configs[1]): a step is one pass of `Segmentation::process` over a batch of B synthetic 1024x1024
RGBA images (uniform noise).  `value` = images/s with the inputs already resident in HBM; `e2e` =
the same metric through the reference-facing C ABI with pinned HOST buffers (H2D of the images and D2H of every image
embedding inside the timed region).  `roofline` = the tcgen05 GEMM family against the measured dense bf16 rate.

Further objects on the same line:
  decoder   config 3: prompt sweep P in {1, 16, 64, 256} on one cached embedding, three output extents, masks/s
  prepost   config 4: 3840x2160 RGB / BGRA / strided resize and the 4K mask upsample, each alone, as HBM GB/s
  latency   config 1: single-call latency of process / compute_mask / compute_masks through the 13-slot table on
            tests/golden/truck.jpg (the reference's fixture) next to README.md:35's figures
  scaleout  config 5: 4096 images sharded over the ranks, process + 16 point prompts per image, end to end
  sustained >= 3 s of back-to-back encoder steps with the clock record
  pcie      pinned-memory copy rates of this rank (alone / both directions at once)

N > 1: one process per GPU (torchrun), images sharded, no data-path collective; only the timing max and the IoU gather
go through NCCL.  `--impl reference` times the CPU stand-in for the reference's ORT path (the PyTorch fp32 oracle:
onnxruntime and the .onnx files are not available offline) on rank 0's host cores, on the same `config`.
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENCODER_GFLOP_PER_IMAGE = 77.54  # SURVEY A.6 (graph as executed, window padding counted)
DECODER_GFLOP_PER_PROMPT = 3.62
README_LATENCY = {"process_ms": {"cpu": 500, "rtx4070": 50}, "compute_mask_ms": {"cpu": 80, "rtx4070": 12},
                  "source": "reference README.md:35 (input resolution and CPU not stated)"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="images per step per GPU")
    ap.add_argument("--prompts", type=int, default=64, help="prompts per decoder step of the headline decoder figure")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=6, help="images timed for the cpu_baseline")
    ap.add_argument("--quick", action="store_true", help="device-resident headline legs only (used for the ncu launch list)")
    ap.add_argument("--job-images", type=int, default=4096, help="images of the config-5 job (all ranks together)")
    ap.add_argument("--sustained", type=float, default=3.0, help="seconds of the sustained encoder run (0 = skip)")
    ap.add_argument("--only", default="", help="comma list of extra sections to run (decoder,prepost,latency,scaleout,sustained,pcie); default all")
    return ap.parse_args()


def bench_config(batch: int, world: int):
    """The `config` object: identical for both arms (the driver compares them)."""
    return {"workload": "MobileSAM Segmentation::process on synthetic 1024x1024 RGBA images (BASELINE configs[1])",
            "images_per_step_per_gpu": batch, "weights": "seeded synthetic MobileSAM (no checkpoint offline)",
            "l2": "inputs cycle through distinct batches totalling > 126 MB (L2)",
            "parallelism": f"image-sharded x{world}, no data-path collective"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []   # (arrival time, text)
        self.t_mark = None

    def mark(self):
        """The timed region starts now: samples that arrived earlier (warm-up) are dropped, unless nothing else came."""
        self.t_mark = time.time()
        return self

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = [l for t, l in self.lines if self.t_mark is None or t >= self.t_mark]
        if not lines and self.lines:  # a very short region between two samples: the one taken just before it
            lines = [self.lines[-1][1]]
        for l in lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


def bind_to_gpu_numa(local_rank: int, world: int):
    """Pins this rank (and every thread / pinned allocation it makes from here on: first touch) to the CPUs next to its GPU.
    When several ranks share one CPU list (a single NUMA node), each takes its own slice of it."""
    info = {"bound": False}
    try:
        bus = subprocess.check_output(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                      text=True).strip().lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(base + "/numa_node").read().strip())
        cpus = []
        for part in open(base + "/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        info.update({"numa_node": node, "gpu_local_cpus": len(allowed)})
        if allowed:
            if world > 1:
                per = max(1, len(allowed) // world)
                mine = allowed[(local_rank * per) % len(allowed):][:per] or allowed
            else:
                mine = allowed
            os.sched_setaffinity(0, mine)
            info.update({"bound": True, "cpus": f"{mine[0]}-{mine[-1]}", "n_cpus": len(mine)})
    except Exception as e:  # containers without sysfs / nvidia-smi: run unbound
        info["error"] = str(e)[:80]
    return info


def oracle_encoder_images_per_s(n_images: int, threads: int, warmup: int = 1):
    """CPU stand-in for the reference's ORT-CPU `process`: fp32 PyTorch oracle, same synthetic weights."""
    import numpy as np
    import torch
    from oracle.mobile_sam_ref import EncoderWithPreprocess, build_synthetic
    torch.set_num_threads(threads)
    enc = EncoderWithPreprocess(build_synthetic(0).image_encoder)
    rng = np.random.default_rng(0)
    imgs = [torch.from_numpy(rng.integers(0, 256, (1024, 1024, 4), dtype=np.uint8)[..., :3].astype(np.float32)) for _ in range(2)]
    with torch.no_grad():
        for _ in range(max(1, warmup)):
            enc(imgs[0])
        t0 = time.perf_counter()
        for i in range(n_images):
            enc(imgs[i % 2])
        dt = time.perf_counter() - t0
    return n_images / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 2  # images per step (bounded sample of the B-image batch)
    v, dt = oracle_encoder_images_per_s(args.steps * sample, threads, warmup=max(1, args.warmup))
    line = {"impl": "reference", "metric": "encoder_images_per_s", "value": v, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args.batch, args.gpus),
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} images/step x {args.steps} steps of the same workload; PyTorch fp32 oracle of the "
                                       "reference's ORT-CPU path (onnxruntime + .onnx models unavailable offline), synthetic weights; "
                                       "one host process whatever --gpus says"},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    binding = bind_to_gpu_numa(local_rank, world)  # before torch spawns threads and before any pinned allocation

    import numpy as np
    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    os.environ["DLIMG_B200_DEVICE"] = str(local_rank)
    os.environ.setdefault("DLIMG_B200_MAX_BATCH", str(args.batch))
    os.environ.setdefault("DLIMG_B200_MAX_PROMPTS", "256")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import dlimgedit_b200 as dl
    from dlimgedit_b200 import synthetic_weights

    sections = set(s for s in args.only.split(",") if s) or {"decoder", "prepost", "latency", "scaleout", "sustained", "pcie"}
    if args.quick:
        sections = set()

    model_dir = tempfile.mkdtemp(prefix=f"dlimg_models_r{rank}_")
    synthetic_weights.write_model_dir(model_dir, seed=0)
    env = dl.Environment(dl.Options(dl.Backend.gpu, model_dir))
    # All library work and the timing events share one explicit (non-default) stream: the legacy default
    # stream has handle 0, which the C ABI reads as "use the environment's own stream".
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    env.set_stream(stream.cuda_stream)

    B, K, W = args.batch, args.steps, args.warmup
    assert W >= 3 or args.quick, "timing rules: at least 3 warm-up steps"
    img_bytes = 1024 * 1024 * 4
    # distinct input batches cycling through > L2 (126 MB) of pixels; per-step activations are ~120 MB/image anyway
    n_sets = max(2, -(-192 * 1024 * 1024 // (B * img_bytes)))
    rng = np.random.default_rng(1000 + rank)  # rank shards are different images (config 5: i % n_gpu sharding)
    host_sets = [torch.from_numpy(rng.integers(0, 256, (B, 1024, 1024, 4), dtype=np.uint8)).pin_memory() for _ in range(n_sets)]
    dev_sets = [h.cuda() for h in host_sets]
    ext = dl.Extent(1024, 1024)

    def dev_views(s):
        return [dl.ImageView(dev_sets[s][i].data_ptr(), ext, dl.Channels.rgba, device=True) for i in range(B)]

    def host_views(s):
        return [dl.ImageView(host_sets[s][i].numpy(), ext, dl.Channels.rgba) for i in range(B)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def timed(fn, steps, sync=None):
        """CUDA events on the work stream around `steps` calls, barrier + synchronize on both sides, max over ranks.
        `sync`: called before the closing event (end-to-end legs: every queued copy has landed)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(i)
        if sync is not None:
            sync()
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---------------- encoder, inputs resident in HBM ----------------
    # Segmentation handles of the last two steps stay alive, older ones are released like a caller would: the
    # embedding stores then recycle through the stream-ordered pool instead of growing it per step.
    keep = collections.deque(maxlen=2)

    def step_dev(i):
        keep.append(env.process_batch(dev_views(i % n_sets)))

    clocks = ClockSampler(local_rank).start()  # (started before the warm-up: nvidia-smi needs ~0.3 s to deliver its first line)
    for i in range(W):
        step_dev(i)
    torch.cuda.synchronize()
    keep.clear()
    clocks.mark()
    launches0 = env.stats()["kernel_launches"]
    ms = timed(step_dev, K)
    launches = env.stats()["kernel_launches"] - launches0
    clock_info = clocks.stop()
    keep.clear()
    value = world * B * K / (ms * 1e-3)

    # ---------------- encoder end to end: pinned host pixels in, embeddings out ----------------
    # Every step uploads its B images from pinned host memory (process_batch on host views: one copy on the library's
    # upload stream) and reads back every embedding (get_embedding_async: the library's download stream).  Nothing waits
    # inside a step, so the upload of step i+1 and the download of step i-1 overlap the encoder of step i; the timed
    # region ends after Environment.synchronize(), when the last embedding has reached the host.
    n_out = 3
    emb_host = [torch.empty(B, 256, 64, 64, dtype=torch.float32).pin_memory() for _ in range(n_out)]

    def step_e2e(i):
        segs = env.process_batch(host_views(i % n_sets))
        dst = emb_host[i % n_out]
        for j, s in enumerate(segs):
            # D2H of the (1,256,64,64) fp32 embedding: the reference keeps it in host memory (segmentation.cpp:124)
            s.embedding_async(dst[j].numpy())
        keep.append(segs)

    k_e2e = max(3, K)
    if args.quick:
        ms_e2e = float("nan")
    else:
        for i in range(min(W, 3)):
            step_e2e(i)
        env.synchronize()
        keep.clear()
        ms_e2e = timed(step_e2e, k_e2e, sync=env.synchronize)
        keep.clear()
    e2e_value = world * B * k_e2e / (ms_e2e * 1e-3)

    # The same loop without the embedding download: what the drop-in's process() does for a GPU caller (the handle keeps the
    # embedding on the device and nothing is read back).  Beside `e2e` it tells how much of the
    # end-to-end time is the 4 MiB/image fp32 download -- at N = 8 the host's aggregate copy bandwidth (see `pcie`).

    def step_e2e_resident(i):
        segs = env.process_batch(host_views(i % n_sets))
        keep.append(segs)

    if args.quick:
        ms_res = float("nan")
    else:
        step_e2e_resident(0)
        env.synchronize()
        keep.clear()
        ms_res = timed(step_e2e_resident, k_e2e, sync=env.synchronize)
        keep.clear()
    # ... and with the embeddings downloaded as fp16 (get_embedding_f16_async, half the D2H bytes)
    emb_host16 = [torch.empty(B, 256, 64, 64, dtype=torch.float16).pin_memory() for _ in range(n_out)]

    def step_e2e_f16(i):
        segs = env.process_batch(host_views(i % n_sets))
        dst = emb_host16[i % n_out]
        for j, s in enumerate(segs):
            s.embedding_f16_async(dst[j].numpy())
        keep.append(segs)

    if args.quick:
        ms_f16 = float("nan")
    else:
        step_e2e_f16(0)
        env.synchronize()
        keep.clear()
        ms_f16 = timed(step_e2e_f16, k_e2e, sync=env.synchronize)
        keep.clear()
    e2e_f16 = {"value": world * B * k_e2e / (ms_f16 * 1e-3), "unit": "images/s", "ms_per_step": ms_f16 / k_e2e,
               "h2d_bytes_per_step": B * 1024 * 1024 * 4, "d2h_bytes_per_step": B * 256 * 64 * 64 * 2,
               "path": "process_batch(host views) + get_embedding_f16_async per image"}
    del emb_host16
    e2e_resident = {"value": world * B * k_e2e / (ms_res * 1e-3), "unit": "images/s", "ms_per_step": ms_res / k_e2e,
                    "h2d_bytes_per_step": B * 1024 * 1024 * 4, "d2h_bytes_per_step": 0,
                    "path": "process_batch(host views) only: embeddings stay on the device behind their handles"}

    # ---------------- decoder: P point prompts on one cached embedding (config 3) ----------------
    P = args.prompts
    seg = env.process_batch(dev_views(0)[:1])[0]
    prng = np.random.default_rng(1)

    def dec_leg(seg_, prompts, multi, steps, warm):
        """masks/s with prompts from the host and masks left on the device."""
        n = 3 if multi else 1
        e = seg_.extent()
        cnt = len(prompts)
        d_masks = torch.empty(cnt, n, e.height, e.width, dtype=torch.uint8, device="cuda")
        d_ious = torch.empty(cnt, n, dtype=torch.float32, device="cuda")
        ptrs = [d_masks[i].data_ptr() for i in range(cnt)]

        def step(i):
            env.compute_masks_batch([seg_] * cnt, prompts, multi=multi, masks_out=ptrs, ious_out=d_ious.data_ptr())

        for i in range(warm):
            step(i)
        t = timed(step, steps)
        return world * cnt * steps / (t * 1e-3), t / steps, d_ious

    prompts = [dl.Point(int(prng.integers(0, 1024)), int(prng.integers(0, 1024))) for _ in range(P)]
    k_dec = max(5, K)
    masks_per_s, ms_dec_step, d_ious = dec_leg(seg, prompts, False, k_dec, max(3, W))

    # end to end: prompts from the host, masks + scores into page-locked host arrays (the reference API hands the caller
    # host masks).  Like the encoder leg: the calls queue their work (host_async), the download of one step runs on the
    # copy-out stream under the decoder of the next one, and the timed region ends after `synchronize`, when every mask of
    # every timed step is in host memory.  Three sets of host buffers, as a pipelined caller would hold.
    masks_per_s_e2e = float("nan")
    if not args.quick:
        n_hset = 3
        h_masks = [torch.empty(P, 1, 1024, 1024, dtype=torch.uint8).pin_memory() for _ in range(n_hset)]
        h_ious = [torch.empty(P, 1, dtype=torch.float32).pin_memory() for _ in range(n_hset)]
        h_list = [[h_masks[k][i].numpy() for i in range(P)] for k in range(n_hset)]

        def step_dec_e2e(i):
            env.compute_masks_batch([seg] * P, prompts, multi=False, host_out=h_list[i % n_hset], host_ious=h_ious[i % n_hset].numpy(),
                                    host_async=True)

        for i in range(3):
            step_dec_e2e(i)
        env.synchronize()
        k_dec_e2e = max(3, k_dec)
        ms_dec_e2e = timed(step_dec_e2e, k_dec_e2e, sync=env.synchronize)
        masks_per_s_e2e = world * P * k_dec_e2e / (ms_dec_e2e * 1e-3)
        del h_masks, h_list

    # ---------------- attribution pass: CUDA events around every kernel launch ----------------
    peaks = load_peaks()
    prof_steps = 2
    prof, prof_dec = {}, {}
    if not args.quick:
        env.profile_enable(True)
        for i in range(prof_steps):
            step_dev(i)
        torch.cuda.synchronize()
        prof = env.profile_read()
        keep.clear()
        d_masks = torch.empty(P, 1024, 1024, dtype=torch.uint8, device="cuda")
        ptrs = [d_masks[i].data_ptr() for i in range(P)]
        for i in range(2):
            env.compute_masks_batch([seg] * P, prompts, multi=False, masks_out=ptrs, ious_out=d_ious.data_ptr())
        torch.cuda.synchronize()
        prof_dec = env.profile_read()
        env.profile_enable(False)
        del d_masks

    total_ms = sum(v["ms"] for v in prof.values()) or 1.0
    kernels = {k: {"ms_per_step": v["ms"] / prof_steps, "launches_per_step": v["launches"] // prof_steps,
                   "share": v["ms"] / total_ms} for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    g = prof.get("gemm_tcgen05_f16", {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
    gemm_tflops = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    # the contraction kernels are bounded by the tensor pipe.  Denominator: the burst figure, because the
    # timed region is far below a second (the GPU stays at its boost clock); the sustained figure is reported beside it.
    timed_s = ms * 1e-3
    peak_tf = peaks["tflops_burst"] if timed_s < 1.0 else peaks["tflops_sustained"]
    launches_g = max(1, g["launches"])
    dram_gbs = (traffic * g["launches"] / (g["ms"] * 1e-3) / 1e9) if (traffic and g["ms"]) else None
    whole_tflops = ENCODER_GFLOP_PER_IMAGE * 1e-3 * B / (ms / K * 1e-3)
    roofline = {"bound": "tensor",
                "kernel": "gemm_tc_kernel<f16> + mlp_fused_kernel (tcgen05.mma + TMA: all encoder GEMMs, the neck 3x3 conv and the fused MLPs)",
                "achieved": gemm_tflops, "peak": peak_tf, "unit": "TFLOP/s", "frac": gemm_tflops / peak_tf,
                "peak_source": peaks["source"] + (", dense bf16 burst (timed region < 1 s)" if timed_s < 1.0 else ", dense bf16 sustained"),
                "flops_per_launch": g["flops"] / launches_g, "ms_per_launch": g["ms"] / launches_g,
                "traffic": traffic, "traffic_source": "profiles/gemm_traffic.json (ncu dram__bytes_read+write per launch)" if traffic else None,
                "frac_of_sustained": gemm_tflops / peaks["tflops_sustained"],
                "dram_gbs_from_measured_traffic": dram_gbs, "dram_frac_of_hbm": (dram_gbs / peaks["hbm_gbs"]) if dram_gbs else None,
                "gemm_flops_per_step": g["flops"] / prof_steps, "gemm_ms_per_step": g["ms"] / prof_steps,
                "gemm_launches_per_step": g["launches"] // prof_steps, "gemm_share_of_step": g["ms"] / total_ms,
                "whole_encoder": {"gflop_per_image": ENCODER_GFLOP_PER_IMAGE, "tflops": whole_tflops, "frac": whole_tflops / peak_tf}}
    total_dec = sum(v["ms"] for v in prof_dec.values()) or 1.0
    dec_kernels = {k: {"ms_per_step": v["ms"] / 2, "share": v["ms"] / total_dec}
                   for k, v in sorted(prof_dec.items(), key=lambda kv: -kv[1]["ms"])}
    mp = prof_dec.get("mask_postprocess")
    mask_post_gbs = (mp["bytes"] / (mp["ms"] * 1e-3) / 1e9) if mp and mp["ms"] else None

    decoder = {"metric": "masks_per_s", "value": masks_per_s, "unit": "masks/s", "prompts_per_step": P,
               "ms_per_step": ms_dec_step, "us_per_prompt": ms_dec_step * 1e3 / P, "mask_extent": "1024x1024",
               "mode": "single mask, point prompts, one cached embedding",
               "e2e": {"value": masks_per_s_e2e, "unit": "masks/s", "d2h_bytes_per_step": P * (1024 * 1024 + 4),
                       "path": "ctypes -> dlimg_b200_Ext.compute_masks_batch(host masks, asynchronous), synchronize at the end"},
               "tflops": DECODER_GFLOP_PER_PROMPT * 1e-3 * P / (ms_dec_step * 1e-3),
               "frac_of_tensor_peak": DECODER_GFLOP_PER_PROMPT * 1e-3 * P / (ms_dec_step * 1e-3) / peaks["tflops_burst"],
               "mask_postprocess_gbs": mask_post_gbs, "mask_postprocess_frac_of_hbm": (mask_post_gbs / peaks["hbm_gbs"]) if mask_post_gbs else None,
               "kernels": dec_kernels}

    # ---------------- config 3: prompt sweep ----------------
    if "decoder" in sections:
        sweep = {}
        for p_cnt in (1, 16, 64, 256):
            pr = [dl.Point(int(prng.integers(0, 1024)), int(prng.integers(0, 1024))) for _ in range(p_cnt)]
            v, t_step, _ = dec_leg(seg, pr, False, max(5, 200 // p_cnt), 3)
            sweep[str(p_cnt)] = {"masks_per_s": v, "ms_per_call": t_step}
        # regions + three-mask mode at 64 prompts
        rg = []
        for _ in range(64):
            x0, y0 = int(prng.integers(0, 1000)), int(prng.integers(0, 1000))
            rg.append(dl.Region(dl.Point(x0, y0), dl.Point(min(1023, x0 + 16 + int(prng.integers(0, 600))), min(1023, y0 + 16 + int(prng.integers(0, 600))))))
        v, t_step, _ = dec_leg(seg, rg, False, 5, 3)
        sweep["64_regions"] = {"masks_per_s": v, "ms_per_call": t_step}
        v, t_step, _ = dec_leg(seg, prompts[:64], True, 5, 3)
        sweep["64_points_3masks"] = {"masks_per_s": v, "ms_per_call": t_step, "note": "three masks per prompt: value counts masks"}
        sweep["64_points_3masks"]["masks_per_s"] = v * 3
        decoder["prompt_sweep_1024"] = sweep
        extents = {}
        for (w_, h_) in ((1800, 1200), (3840, 2160)):
            img = torch.from_numpy(np.random.default_rng(2).integers(0, 256, (h_, w_, 3), dtype=np.uint8)).cuda()
            sg = env.process_batch([dl.ImageView(img.data_ptr(), dl.Extent(w_, h_), dl.Channels.rgb, device=True)])[0]
            pr = [dl.Point(int(prng.integers(0, w_)), int(prng.integers(0, h_))) for _ in range(16)]
            v, t_step, _ = dec_leg(sg, pr, False, 5, 3)
            extents[f"{w_}x{h_}"] = {"prompts": 16, "masks_per_s": v, "ms_per_call": t_step, "mask_mb": w_ * h_ / 1e6}
            sg.close()
            del img
        decoder["extents"] = extents

    # ---------------- config 4: 4K pre- / post-processing, each stage alone ----------------
    prepost = None
    if "prepost" in sections:
        import ctypes
        prepost = {"peak_hbm_gbs": peaks["hbm_gbs"], "note": "stand-alone stages through dlimg_b200_Ext, device buffers, CUDA events; "
                   "bytes = SURVEY 8(d) algorithmic bytes (input read once + output written once)"}
        w_, h_ = 3840, 2160
        r2 = np.random.default_rng(2)

        def time_resize(ch, bpp, stride):
            n_in = 6  # 6 distinct inputs of 25-33 MB: > L2 between repeats
            bufs = [torch.from_numpy(r2.integers(0, 256, (h_, stride), dtype=np.uint8)).cuda() for _ in range(n_in)]
            out = torch.empty(576 * 1024 * bpp, dtype=torch.uint8, device="cuda")
            ext2 = (ctypes.c_int * 2)()

            def step(i):
                v = dl.ImageView(bufs[i % n_in].data_ptr(), dl.Extent(w_, h_), ch, stride, device=True).to_c()
                r = dl.ext().resize_longest_side(env.handle(), ctypes.byref(v), 1024, out.data_ptr(), ext2)
                assert r == 0, dl.api().last_error()

            for i in range(3):
                step(i)
            t = timed(step, 24) / 24
            nbytes = h_ * w_ * bpp + 1024 * 576 * bpp
            return {"ms": t, "gbs": nbytes / (t * 1e-3) / 1e9, "frac_of_hbm": nbytes / (t * 1e-3) / 1e9 / peaks["hbm_gbs"],
                    "bytes": nbytes, "out_extent": [ext2[0], ext2[1]]}

        prepost["resize_3840x2160_rgb"] = time_resize(dl.Channels.rgb, 3, w_ * 3)
        prepost["resize_3840x2160_bgra"] = time_resize(dl.Channels.bgra, 4, w_ * 4)
        prepost["resize_3840x2160_rgb_strided"] = time_resize(dl.Channels.rgb, 3, w_ * 3 + 64)
        # mask upsample + threshold to 4K (16 planes per launch) and to 1024^2 (64 planes: 16 would be 7 us, launch-bound)
        for (mw, mh, cnt) in ((3840, 2160, 16), (1024, 1024, 64)):
            n_in = 8
            lows = [torch.randn(cnt, 256, 256, device="cuda") for _ in range(n_in)]
            outm = torch.empty(cnt, mh, mw, dtype=torch.uint8, device="cuda")

            def step(i, lows=lows, outm=outm, mw=mw, mh=mh, cnt=cnt):
                r = dl.ext().mask_postprocess(env.handle(), lows[i % len(lows)].data_ptr(), cnt, mw, mh, outm.data_ptr())
                assert r == 0, dl.api().last_error()

            for i in range(3):
                step(i)
            t = timed(step, 24) / 24
            nbytes = cnt * (256 * 256 * 4 + mw * mh)
            prepost[f"mask_upsample_{mw}x{mh}"] = {"planes": cnt, "ms": t, "gbs": nbytes / (t * 1e-3) / 1e9,
                                                  "frac_of_hbm": nbytes / (t * 1e-3) / 1e9 / peaks["hbm_gbs"], "bytes": nbytes}
            del lows, outm

    # ---------------- config 1: single-call latency through the 13-slot table, real image ----------------
    latency = None
    if "latency" in sections and rank == 0:
        truck = os.path.join(ROOT, "tests", "golden", "truck.jpg")
        # like the reference's C++ wrapper: the input is an Image of the library (Image::load -> load_image, the library's own
        # JPEG reader) and the masks are Images it creates (create_image) -- page-locked while the environment lives
        timg = dl.Image.load(truck)
        view = timg.view()
        view_pageable = dl.ImageView(timg.pixels.copy(), channels=dl.Channels.rgb)  # a caller's own (pageable) pixels

        def wall(fn, n):
            fn()  # first call (graph capture / allocations) is reported separately by the caller when wanted
            ts = []
            for _ in range(n):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                fn()
                ts.append((time.perf_counter() - t0) * 1e3)
            return {"median_ms": statistics.median(ts), "min_ms": min(ts), "n": n}

        import ctypes
        plain_mask = np.empty((timg.extent().height, timg.extent().width), np.uint8)

        def raw_mask(seg, x, y):  # the slot itself, writing into the caller's own (pageable) buffer
            ptrs = (ctypes.c_void_p * 3)(plain_mask.ctypes.data, None, None)
            acc = (ctypes.c_float * 3)()
            assert dl.api().get_segmentation_mask(seg._h, (ctypes.c_int * 2)(x, y), None, ptrs, acc) == 0

        env.synchronize()
        t0 = time.perf_counter()
        s0 = dl.Segmentation.process(view, env)
        env.synchronize()
        first_ms = (time.perf_counter() - t0) * 1e3
        segs_l = []
        lat = {"image": "tests/golden/truck.jpg (reference test/input/truck.jpg, 1800x1200 RGB, decoded by the library's load_image)",
               "api": "dlimg_Api 13-slot table: process_image_for_segmentation / get_segmentation_mask (blocking, host buffers "
                      "from load_image / create_image as in dlimgedit.impl.hpp; *_pageable: the caller's own numpy pixels)",
               "process_first_call_ms": first_ms,
               # Segmentation::process returns once the pixels are consumed (the encoder keeps running); the reference's call
               # returns with the embedding computed, so the latency that compares is "call + synchronize"
               "process": wall(lambda: (segs_l.append(dl.Segmentation.process(view, env)), env.synchronize()), 10),
               "process_pageable": wall(lambda: (segs_l.append(dl.Segmentation.process(view_pageable, env)), env.synchronize()), 10),
               "process_call_returns_ms": wall(lambda: segs_l.append(dl.Segmentation.process(view, env)), 5)["median_ms"],
               "compute_mask_point_486_722": wall(lambda: s0.compute_mask(dl.Point(486, 722)), 20),  # test_segmentation.cpp:139
               "compute_mask_point_486_722_pageable": wall(lambda: raw_mask(s0, 486, 722), 20),
               "compute_mask_point_220_355": wall(lambda: s0.compute_mask(dl.Point(220, 355)), 20),  # README.md:29
               "compute_mask_region": wall(lambda: s0.compute_mask(dl.Region(dl.Point(180, 110), dl.Point(505, 330))), 20),
               "compute_masks_point": wall(lambda: s0.compute_masks(dl.Point(486, 722)), 20),
               "reference_readme": README_LATENCY}
        m = s0.compute_mask(dl.Point(486, 722))
        lat["mask_486_722_coverage"] = float((m > 0).mean())
        for s_ in segs_l:
            s_.close()
        s0.close()
        latency = lat

    # ---------------- sustained run (>= 3 s) with its clock record ----------------
    sustained = None
    if "sustained" in sections and args.sustained > 0:
        n_steps = max(K, int(args.sustained / (ms / K * 1e-3)) + 1)
        cs = ClockSampler(local_rank).start()
        t = timed(step_dev, n_steps)
        ci = cs.stop()
        keep.clear()
        sustained = {"seconds": t * 1e-3, "steps": n_steps, "images_per_s": world * B * n_steps / (t * 1e-3), "clocks": ci,
                     "vs_headline": world * B * n_steps / (t * 1e-3) / value}

    # ---------------- pinned-memory copy rates of this rank ----------------
    pcie = None
    if "pcie" in sections:
        nb = 256 << 20
        hsrc, hdst = torch.empty(nb, dtype=torch.uint8).pin_memory(), torch.empty(nb, dtype=torch.uint8).pin_memory()
        dsrc, ddst = torch.empty(nb, dtype=torch.uint8, device="cuda"), torch.empty(nb, dtype=torch.uint8, device="cuda")
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

        def copy_rate(h2d, d2h, reps=4):
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                if h2d:
                    with torch.cuda.stream(s_in):
                        ddst.copy_(hsrc, non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s_out):
                        hdst.copy_(dsrc, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            dt = max_over_ranks(dt * 1e3) * 1e-3
            return nb * reps / dt / 1e9

        copy_rate(True, True, 1)
        pcie = {"h2d_gbs": copy_rate(True, False), "d2h_gbs": copy_rate(False, True), "both_gbs_each": copy_rate(True, True),
                "note": "per rank, all ranks copying at once (max time over ranks); 256 MiB pinned buffers",
                "e2e_needs_gbs_each": e2e_value / world * img_bytes / 1e9 if e2e_value == e2e_value else None}
        del hsrc, hdst, dsrc, ddst

    # ---------------- config 5: 4096 images x 16 prompts, sharded, end to end ----------------
    scaleout = None
    if "scaleout" in sections:
        per_rank = args.job_images // world
        n_steps = max(1, per_rank // B)
        ppi = 16
        job_prng = np.random.default_rng(4 + rank)
        job_prompts = [[dl.Point(int(job_prng.integers(0, 1024)), int(job_prng.integers(0, 1024))) for _ in range(B * ppi)] for _ in range(2)]
        d_masks = torch.empty(B * ppi, 1024, 1024, dtype=torch.uint8, device="cuda")  # masks stay on the device (optional in config 5)
        mptrs = [d_masks[i].data_ptr() for i in range(B * ppi)]
        d_iou = torch.empty(n_steps, B * ppi, dtype=torch.float32, device="cuda")
        h_iou = torch.empty(n_steps, B * ppi, dtype=torch.float32).pin_memory()

        def job_step(i):
            segs = env.process_batch(host_views(i % n_sets))          # H2D of the step's 32 images inside
            owners = [s for s in segs for _ in range(ppi)]            # 16 prompts per image, 512 per step
            env.compute_masks_batch(owners, job_prompts[i % 2], multi=False, masks_out=mptrs, ious_out=d_iou[i].data_ptr())
            for s in segs:
                s.close()

        for i in range(2):
            job_step(i)
        torch.cuda.synchronize()

        def job_sync():
            h_iou.copy_(d_iou, non_blocking=True)  # the job's result: one predicted IoU per (image, prompt)
            env.synchronize()
            torch.cuda.synchronize()

        cs = ClockSampler(local_rank).start()
        t = timed(job_step, n_steps, sync=job_sync)
        ci = cs.stop()
        images = world * n_steps * B
        # config 5's one exchange: (P x IoU f32) per image to every rank (image i lives on rank i mod N) -- tiny, after
        # the timed region; the same helper the world_size-2 gloo test covers (tests/test_multi_rank_cpu.py)
        from dlimgedit_b200 import sharding
        full = sharding.gather_scores(h_iou.numpy().reshape(n_steps * B, ppi), images, rank, world)
        scaleout = {"job": f"{args.job_images} images of 1024x1024 RGBA sharded over {world} rank(s), process + {ppi} point prompts per image "
                           "(single mask each), end to end: pinned host pixels in, IoU scores out, masks left on the device",
                    "images": images, "prompts": images * ppi, "seconds": t * 1e-3, "images_per_s": images / (t * 1e-3),
                    "masks_per_s": images * ppi / (t * 1e-3), "steps_per_rank": n_steps, "clocks": ci,
                    "h2d_bytes_per_step": B * img_bytes, "d2h_bytes_total": int(n_steps * B * ppi * 4),
                    "gathered_scores_shape": list(full.shape), "gather": "sharding.gather_scores (all_gather over NCCL when N > 1)"}
        del d_masks

    # ---------------- CPU baseline (rank 0, N == 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.quick:
        threads = os.cpu_count() or 1
        v, _ = oracle_encoder_images_per_s(args.cpu_sample, threads)
        cpu = {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_sample} images of the same workload; PyTorch fp32 oracle standing in for the reference's "
                         "ORT-CPU path (onnxruntime/.onnx unavailable offline); README.md:35 quotes ~2 images/s on an unnamed CPU"}

    if rank == 0:
        line = {
            "metric": "encoder_images_per_s", "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 storage / f32 accumulate (tcgen05 kind::f16)", "data": "synthetic",
            "config": bench_config(B, world),
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * img_bytes,
                    "d2h_bytes_per_step": B * 256 * 64 * 64 * 4, "ms_per_step": ms_e2e / k_e2e,
                    "path": "ctypes -> dlimg_b200_Ext.process_batch(host views) + get_embedding_async per image, synchronize at the end"},
            "e2e_resident": e2e_resident,
            "e2e_f16_download": e2e_f16,
            "gpu_launches": int(launches),
            "clocks": clock_info,
            "roofline": roofline,
            "kernels": kernels,
            "decoder": decoder,
            "host_binding": binding,
        }
        for k_, v_ in (("prepost", prepost), ("latency", latency), ("scaleout", scaleout), ("sustained", sustained), ("pcie", pcie),
                       ("cpu_baseline", cpu)):
            if v_ is not None:
                line[k_] = v_
        print(json.dumps(line))
    seg.close()
    env.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
