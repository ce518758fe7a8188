"""One eager decoder pass (64 prompts) from an ncu launch list: time, DRAM bytes and rate per launch.
    python tools/decoder_table.py profiles/r02_launches_decoder.csv >> profiles/r02_roofline_table.md"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    HBM = 6537.6
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
L = collections.OrderedDict()
for r in rows[hdr + 1:]:
    d = L.setdefault(int(r[0]), {"name": r[4], "grid": r[8]})
    d[r[12]] = float(r[14].replace(",", ""))
starts = [i for i in L if "prompt_tokens_kernel" in L[i]["name"]]
s0 = starts[-1]
ids = [i for i in L if i >= s0 and ("dlimg" in L[i]["name"] or "gemm" in L[i]["name"] or "prepost" in L[i]["name"] or "dec::" in L[i]["name"])]
ids = [i for i in ids if "at::" not in L[i]["name"]]
tot = sum(L[i]["gpu__time_duration.sum"] for i in ids) / 1000.0
tb = sum(L[i].get("dram__bytes_read.sum", 0) + L[i].get("dram__bytes_write.sum", 0) for i in ids)
print(f"\n# One eager decoder pass, 64 point prompts on one embedding, 1024^2 masks\n")
print(f"Source: `{os.path.basename(sys.argv[1])}` (ncu, graphs off, serialised: {tot:.0f} us; 712-743 us as a CUDA graph).  `tools/decoder_table.py`.\n")
print("| # | kernel | grid | us | DRAM read MB | written MB | GB/s | of HBM |")
print("|---|---|---|---|---|---|---|---|")
for k, i in enumerate(ids):
    d = L[i]
    us = d["gpu__time_duration.sum"] / 1000.0
    rd, wr = d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6
    name = d["name"].split("(")[0].replace("void ", "").replace("dlimg::", "").replace("<unnamed>::", "")
    gbs = (rd + wr) / us * 1e3
    print(f"| {k} | `{name[:48]}` | {d['grid']} | {us:.1f} | {rd:.1f} | {wr:.1f} | {gbs:.0f} | {gbs / HBM:.2f} |")
print(f"| | **pass** | | {tot:.1f} | | {tb / 1e6:.0f} (both) | {tb / tot / 1e3:.0f} | {tb / tot / 1e3 / HBM:.2f} |")
print("\n(The image-side GEMMs write tensors that the next kernel reads back out of the 126 MB L2, so their DRAM columns undercount what they move:\n"
      "the [K|V|Q] projection of 64 prompts reads 134 MB and writes 201 MB of activations in 100 us = 3.4 TB/s of kernel traffic.)")
