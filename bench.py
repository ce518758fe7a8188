#!/usr/bin/env python
"""bench.py -- MobileSAM segmentation hot path on B200 (BASELINE.json configs[1] + configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--prompts P] [--impl ours|reference]

A step is one pass of `Segmentation::process` over a batch of B synthetic 1024x1024 RGBA images (uniform
noise, numpy default_rng(0)).  `value` = images/s with the inputs already resident in
HBM; `e2e` = the same metric through the reference-facing C ABI with pinned HOST buffers (H2D of the images
and D2H of every image embedding inside the timed region).  A `decoder` object reports masks/s for P point
prompts against one cached embedding (config 3), device-resident and end-to-end.

N > 1: one process per GPU (torchrun), images sharded, no data-path collective; only the timing max and the
IoU gather go through NCCL.  `--impl reference` times the CPU stand-in for the reference's ORT path (the
PyTorch fp32 oracle: onnxruntime and the .onnx files are not available offline) on rank 0's host cores.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="images per step per GPU")
    ap.add_argument("--prompts", type=int, default=64, help="prompts per decoder step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=6, help="images timed for the cpu_baseline")
    ap.add_argument("--quick", action="store_true", help="device-resident legs only (used for the ncu launch list)")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_encoder_images_per_s(n_images: int, threads: int):
    """CPU stand-in for the reference's ORT-CPU `process`: fp32 PyTorch oracle, same synthetic weights."""
    import numpy as np
    import torch
    from oracle.mobile_sam_ref import EncoderWithPreprocess, build_synthetic
    torch.set_num_threads(threads)
    enc = EncoderWithPreprocess(build_synthetic(0).image_encoder)
    rng = np.random.default_rng(0)
    imgs = [torch.from_numpy(rng.integers(0, 256, (1024, 1024, 4), dtype=np.uint8)[..., :3].astype(np.float32)) for _ in range(2)]
    with torch.no_grad():
        enc(imgs[0])  # warm-up
        t0 = time.perf_counter()
        for i in range(n_images):
            enc(imgs[i % 2])
        dt = time.perf_counter() - t0
    return n_images / dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch
    from oracle.mobile_sam_ref import EncoderWithPreprocess, build_synthetic
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    enc = EncoderWithPreprocess(build_synthetic(0).image_encoder)
    sample = 2  # images per step (bounded sample of the B-image batch)
    rng = np.random.default_rng(0)
    imgs = [torch.from_numpy(rng.integers(0, 256, (1024, 1024, 4), dtype=np.uint8)[..., :3].astype(np.float32)) for _ in range(sample)]
    with torch.no_grad():
        for _ in range(max(1, args.warmup)):
            enc(imgs[0])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            for im in imgs:
                enc(im)
        dt = time.perf_counter() - t0
    v = args.steps * sample / dt
    line = {"impl": "reference", "metric": "encoder_images_per_s", "value": v, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "MobileSAM Segmentation::process, synthetic 1024x1024 RGBA images (configs[1])",
                       "images_per_step": sample},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} images/step x {args.steps} steps; PyTorch fp32 oracle of the reference's ORT-CPU path "
                                       "(onnxruntime + .onnx models unavailable offline), synthetic weights"},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    os.environ["DLIMG_B200_DEVICE"] = str(local_rank)
    os.environ.setdefault("DLIMG_B200_MAX_BATCH", str(args.batch))
    os.environ.setdefault("DLIMG_B200_MAX_PROMPTS", "64")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import dlimgedit_b200 as dl
    from dlimgedit_b200 import synthetic_weights

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
    img_bytes = 1024 * 1024 * 4
    # distinct input batches cycling through > L2 (126 MB) of pixels; per-step activations are ~200 MB/image anyway
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

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(i)
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---------------- encoder, inputs resident in HBM ----------------
    # Segmentation handles of the last two steps stay alive, older ones are released like a caller would: the
    # embedding stores then recycle through the stream-ordered pool instead of growing it by 32 MiB per step
    # (pool growth is a 2-6 ms host-side driver call, tools/step_times.py).
    keep = collections.deque(maxlen=2)

    def step_dev(i):
        keep.append(env.process_batch(dev_views(i % n_sets)))

    for i in range(W):
        step_dev(i)
    torch.cuda.synchronize()
    keep.clear()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = env.stats()["kernel_launches"]
    ms = timed(step_dev, K)
    launches = env.stats()["kernel_launches"] - launches0
    clock_info = clocks.stop()
    keep.clear()
    value = world * B * K / (ms * 1e-3)

    # ---------------- encoder end to end: pinned host pixels in, embeddings out ----------------
    # Every step uploads its B images from pinned host memory (process_batch on host views: the copy runs on the
    # library's upload stream) and reads back every embedding (get_embedding_async: the library's download stream).
    # Nothing waits inside a step, so the upload of step i+1 and the download of step i-1 overlap the encoder of
    # step i; the timed region ends after Environment.synchronize(), when the last embedding has reached the host.
    n_out = 3
    emb_host = [torch.empty(B, 256, 64, 64, dtype=torch.float32).pin_memory() for _ in range(n_out)]

    def step_e2e(i):
        segs = env.process_batch(host_views(i % n_sets))
        dst = emb_host[i % n_out]
        for j, s in enumerate(segs):
            # D2H of the (1,256,64,64) fp32 embedding: the reference keeps it in host memory (segmentation.cpp:124)
            s.embedding_async(dst[j].numpy())
        keep.append(segs)

    def timed_e2e(steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            step_e2e(i)
        env.synchronize()  # all uploads, encoders and downloads of the timed steps are complete
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    k_e2e = max(3, K)
    if args.quick:
        ms_e2e = float("nan")
    else:
        for i in range(min(W, 2)):
            step_e2e(i)
        env.synchronize()
        keep.clear()
        ms_e2e = timed_e2e(k_e2e)
        keep.clear()
    e2e_value = world * B * k_e2e / (ms_e2e * 1e-3)

    # ---------------- decoder: P point prompts on one cached embedding (config 3) ----------------
    P = args.prompts
    seg = env.process_batch(dev_views(0)[:1])[0]
    prng = np.random.default_rng(1)
    prompts = [dl.Point(int(prng.integers(0, 1024)), int(prng.integers(0, 1024))) for _ in range(P)]
    d_masks = torch.empty(P, 1024, 1024, dtype=torch.uint8, device="cuda")
    d_ious = torch.empty(P, dtype=torch.float32, device="cuda")
    ptrs = [d_masks[i].data_ptr() for i in range(P)]

    def step_dec(i):
        env.compute_masks_batch([seg] * P, prompts, multi=False, masks_out=ptrs, ious_out=d_ious.data_ptr())

    for i in range(W):
        step_dec(i)
    k_dec = max(5, K)
    ms_dec = timed(step_dec, k_dec)
    masks_per_s = world * P * k_dec / (ms_dec * 1e-3)

    # end to end: prompts from the host, masks + scores into page-locked host arrays (the reference API hands the caller
    # host masks).  Like the encoder leg: the calls queue their work (host_async), the download of one step runs on the
    # copy-out stream under the decoder of the next one, and the timed region ends after `synchronize`, when every mask of
    # every timed step is in host memory.  Three sets of host buffers, as a pipelined caller would hold.
    n_hset = 3
    h_masks = [torch.empty(P, 1, 1024, 1024, dtype=torch.uint8).pin_memory() for _ in range(n_hset)]
    h_ious = [torch.empty(P, 1, dtype=torch.float32).pin_memory() for _ in range(n_hset)]
    h_list = [[h_masks[k][i].numpy() for i in range(P)] for k in range(n_hset)]

    def step_dec_e2e(i):
        env.compute_masks_batch([seg] * P, prompts, multi=False, host_out=h_list[i % n_hset], host_ious=h_ious[i % n_hset].numpy(),
                                host_async=True)

    def timed_dec_e2e(steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            step_dec_e2e(i)
        env.synchronize()  # every decoder pass and every download of the timed steps is complete
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    k_dec_e2e = max(3, k_dec)
    if args.quick:
        ms_dec_e2e = float("nan")
    else:
        step_dec_e2e(0)
        env.synchronize()
        ms_dec_e2e = timed_dec_e2e(k_dec_e2e)
    masks_per_s_e2e = world * P * k_dec_e2e / (ms_dec_e2e * 1e-3)

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
        for i in range(2):
            step_dec(i)
        torch.cuda.synchronize()
        prof_dec = env.profile_read()
        env.profile_enable(False)

    total_ms = sum(v["ms"] for v in prof.values()) or 1.0
    kernels = {k: {"ms_per_step": v["ms"] / prof_steps, "launches_per_step": v["launches"] // prof_steps,
                   "share": v["ms"] / total_ms} for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    gk = "gemm_tcgen05_f16"  # the 16-bit (fp16 storage) tcgen05 GEMM
    g = prof.get(gk, {"ms": 0.0, "flops": 0.0, "launches": 0})
    achieved = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    # The GEMM family of this model has an arithmetic intensity of ~100 flop/byte (1.055 PFLOP over 10.4 GB per 16-image
    # pass), below the B200 ridge point (measured 1407.6 TFLOP/s / 6537.6 GB/s = 215 flop/byte): its roofline is HBM.
    # achieved = algorithmic bytes (A + B + C [+ residual], each once) of its launches / their CUDA-event time.
    gemm_gbs = g.get("bytes", 0.0) / (g["ms"] * 1e-3) / 1e9 if g["ms"] else 0.0
    roofline = {"bound": "hbm", "kernel": "gemm_tc_kernel<f16> + mlp_fused_kernel (tcgen05.mma + TMA: all encoder GEMMs, the neck 3x3 conv and the fused MLPs)", "achieved": gemm_gbs,
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gemm_gbs / peaks["hbm_gbs"],
                "peak_source": peaks["source"] + " STREAM-style copy bandwidth",
                "traffic": traffic, "algorithmic_bytes_per_launch": (g.get("bytes", 0.0) / g["launches"]) if g["launches"] else None,
                "tensor_tflops": achieved, "tensor_frac_of_sustained_bf16": achieved / peaks["tflops_sustained"],
                "gemm_flops_per_step": g["flops"] / prof_steps, "gemm_ms_per_step": g["ms"] / prof_steps,
                "gemm_launches_per_step": g["launches"] // prof_steps, "gemm_share_of_step": g["ms"] / total_ms,
                "whole_encoder_tflops": ENCODER_GFLOP_PER_IMAGE * 1e-3 * B / (ms / K * 1e-3)}
    total_dec = sum(v["ms"] for v in prof_dec.values()) or 1.0
    dec_kernels = {k: {"ms_per_step": v["ms"] / 2, "share": v["ms"] / total_dec}
                   for k, v in sorted(prof_dec.items(), key=lambda kv: -kv[1]["ms"])}

    # ---------------- CPU baseline (rank 0, N == 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.quick:
        threads = os.cpu_count() or 1
        v = oracle_encoder_images_per_s(args.cpu_sample, threads)
        cpu = {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_sample} images of the same workload; PyTorch fp32 oracle standing in for the reference's "
                         "ORT-CPU path (onnxruntime/.onnx unavailable offline); README.md:35 quotes ~2 images/s on an unnamed CPU"}

    if world > 1:  # the one exchange of the sharded job: gather per-prompt IoU scores (tiny)
        out = [torch.empty_like(d_ious) for _ in range(world)]
        dist.all_gather(out, d_ious)

    if rank == 0:
        line = {
            "metric": "encoder_images_per_s", "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 storage / f32 accumulate (tcgen05 kind::f16); decoder tf32/f32", "data": "synthetic",
            "config": {"workload": "MobileSAM Segmentation::process on synthetic 1024x1024 RGBA images (BASELINE configs[1])",
                       "images_per_step_per_gpu": B, "weights": "seeded synthetic MobileSAM (no checkpoint offline)",
                       "l2": f"inputs cycle through {n_sets} distinct batches = {n_sets * B * img_bytes >> 20} MiB (> 126 MB L2)",
                       "parallelism": f"image-sharded x{world}, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * img_bytes,
                    "d2h_bytes_per_step": B * 256 * 64 * 64 * 4, "ms_per_step": ms_e2e / k_e2e,
                    "path": "ctypes -> dlimg_b200_Ext.process_batch(host views) + get_embedding_async per image, synchronize at the end"},
            "gpu_launches": int(launches),
            "clocks": clock_info,
            "roofline": roofline,
            "kernels": kernels,
            "decoder": {"metric": "masks_per_s", "value": masks_per_s, "unit": "masks/s", "prompts_per_step": P,
                        "ms_per_step": ms_dec / k_dec, "mask_extent": "1024x1024", "mode": "single mask, point prompts",
                        "e2e": {"value": masks_per_s_e2e, "unit": "masks/s", "d2h_bytes_per_step": P * (1024 * 1024 + 4),
                                "path": "ctypes -> dlimg_b200_Ext.compute_masks_batch(host masks, asynchronous), synchronize at the end"},
                        "tflops": DECODER_GFLOP_PER_PROMPT * 1e-3 * P / (ms_dec / k_dec * 1e-3), "kernels": dec_kernels},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
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
