import json, sys
d = json.loads([l for l in open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/bench.json').read().splitlines() if l.startswith('{')][-1])  # NCCL prints its version first
print("value", round(d["value"], 1), "img/s  ms/step", round(d["ms_per_step"], 3), " e2e", round(d["e2e"]["value"], 1), "ms", round(d["e2e"]["ms_per_step"], 2))
print("launches", d["gpu_launches"], d["clocks"])
r = d["roofline"]
print("gemm achieved", round(r["achieved"], 1), r["unit"], "frac", round(r["frac"], 3), "gemm ms", round(r["gemm_ms_per_step"], 3), "share", round(r["gemm_share_of_step"], 3),
      "whole enc", {k: round(v, 3) for k, v in r["whole_encoder"].items()})
for k, v in d["kernels"].items():
    print("  ", k, round(v["ms_per_step"], 3), v["launches_per_step"], round(v["share"], 3))
dd = d["decoder"]
print("decoder", round(dd["value"]), "masks/s  ms/step", round(dd["ms_per_step"], 3), " e2e", round(dd["e2e"]["value"]), "mask_post GB/s", dd.get("mask_postprocess_gbs"))
for k, v in dd["kernels"].items():
    print("  ", k, round(v["ms_per_step"], 3), round(v["share"], 3))
for k in ("prompt_sweep_1024", "extents"):
    if k in dd:
        print(" ", k, json.dumps(dd[k]))
for k in ("prepost", "latency", "scaleout", "sustained", "pcie", "host_binding", "cpu_baseline"):
    if k in d:
        print(k, json.dumps(d[k]))
