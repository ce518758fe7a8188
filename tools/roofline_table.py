"""Per-launch roofline table of one encoder pass at batch 32 from an ncu launch list (gpu__time_duration + DRAM bytes).

    python tools/roofline_table.py profiles/r02_launches_bench.csv > profiles/r02_roofline_table.md

FLOPs per launch come from the TinyViT-5M / MobileSAM shapes (B = 32 images of 1024^2), DRAM bytes from
ncu, peaks from MEASURED_PEAKS.json (fallback: the profiling guide's figures).  ncu times are cold-cache and serialised: the
shares, not the absolutes, compare with the live step."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = 32
try:
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    HBM = float(pk.get("hbm_gbs", pk.get("hbm_gbs_burst", 6537.6)))
    TF = float(pk.get("bf16_tflops", pk.get("bf16_tflops_burst", 1649.5)))
except Exception:  # noqa: BLE001
    HBM, TF = 6537.6, 1649.5

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
L = collections.OrderedDict()
for r in rows[hdr + 1:]:
    d = L.setdefault(int(r[0]), {"name": r[4]})
    d[r[12]] = float(r[14].replace(",", ""))
starts = [i for i in L if "patch_embed" in L[i]["name"]]
s0 = starts[1]
e = starts[2] if len(starts) > 2 else max(L) + 1
launches = [L[i] for i in range(s0, e)]

T1, T2 = B * 128 * 128, B * 64 * 64  # tokens at 128^2 and 64^2
PX0 = B * 256 * 256


def gemm(m, k, n):
    return 2.0 * m * k * n


def block(tokens, c, heads, ws, res, hidden_fused):
    nw = -(-res // ws)
    n = ws * ws
    out = [("qkv GEMM (LayerNorm folded)", gemm(tokens, c, 3 * c)),
           (f"window attention {ws}x{ws}", B * nw * nw * heads * 4.0 * n * n * 32),
           ("proj GEMM + residual", gemm(tokens, c, c)),
           ("local_conv 3x3 depthwise + row sums", tokens * c * 18.0)]
    if hidden_fused:
        out.append(("fused MLP (LN, fc1, GELU, fc2, residual)", 2 * gemm(tokens, c, 4 * c)))
    else:
        out += [("fc1 GEMM (LN folded) + GELU", gemm(tokens, c, 4 * c)), ("fc2 GEMM + residual", gemm(tokens, 4 * c, c))]
    return out


plan = [("PatchEmbed: preprocess + conv3x3 s2 + GELU + conv3x3 s2", B * (512 * 512 * 32 * 27 * 2.0 + 256 * 256 * 64 * 288 * 2.0))]
for _ in range(2):
    plan += [("MBConv expand 1x1 64->256 + GELU", gemm(PX0, 64, 256)),
             ("MBConv tail: dw3x3 + GELU + 1x1 256->64 + shortcut + GELU", PX0 * 256 * 18.0 + gemm(PX0, 256, 64))]
plan += [("PatchMerging 1x1 64->128 + GELU", gemm(PX0, 64, 128)), ("PatchMerging dw3x3 s2 + GELU", T1 * 128 * 18.0), ("PatchMerging 1x1 128->128", gemm(T1, 128, 128))]
for _ in range(2):
    plan += block(T1, 128, 4, 7, 128, True)
plan += [("PatchMerging 1x1 128->160 + GELU", gemm(T1, 128, 160)), ("PatchMerging dw3x3 s2 + GELU", T2 * 160 * 18.0), ("PatchMerging 1x1 160->160", gemm(T2, 160, 160))]
for _ in range(6):
    plan += block(T2, 160, 5, 14, 64, True)
plan += [("PatchMerging 1x1 160->320 + GELU", gemm(T2, 160, 320)), ("PatchMerging dw3x3 + GELU", T2 * 320 * 18.0), ("PatchMerging 1x1 320->320", gemm(T2, 320, 320))]
for _ in range(2):
    plan += block(T2, 320, 10, 7, 64, False)
plan += [("neck 1x1 320->256", gemm(T2, 320, 256)), ("neck LayerNorm2d", T2 * 256 * 8.0), ("neck 3x3 256->256 (implicit GEMM)", gemm(T2, 2304, 256)),
         ("neck LayerNorm2d -> NCHW fp32 + 16-bit keys", T2 * 256 * 8.0), ("decoder layer-0 K|V|Q projections of the image", gemm(T2, 256, 384))]
assert len(plan) == len(launches), (len(plan), len(launches))

agg = collections.OrderedDict()
for (what, flops), d in zip(plan, launches):
    a = agg.setdefault(what, [0, 0.0, 0.0, 0.0, d["name"].split("(")[0].split("::")[-1][:40]])
    a[0] += 1
    a[1] += d["gpu__time_duration.sum"] / 1000.0
    a[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    a[3] += flops
tot = sum(a[1] for a in agg.values())
print(f"# One encoder pass, 32 images of 1024^2: every launch against both rooflines\n")
print(f"Source: `{os.path.basename(sys.argv[1])}` (ncu, cold cache, serialised: {tot:.0f} us for the pass; the live step is shorter) and the")
print(f"model's shapes; peaks {HBM:.1f} GB/s and {TF:.1f} TFLOP/s (dense bf16, burst).  `tools/roofline_table.py` regenerates it.\n")
print("| step | kernel | launches | us | share | DRAM MB | GB/s | of HBM | GFLOP | TFLOP/s | of tensor peak |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for what, (cnt, us, byt, fl, kern) in agg.items():
    gbs = byt / us / 1e3
    tfs = fl / us / 1e6
    tensor = "–" if ("LayerNorm2d" in what or "depthwise" in what or "dw3x3 s2" in what or what.endswith("dw3x3 + GELU")) else f"{tfs / TF:.2f}"
    print(f"| {what} | `{kern}` | {cnt} | {us:.1f} | {100 * us / tot:.1f}% | {byt / 1e6:.0f} | {gbs:.0f} | {gbs / HBM:.2f} | {fl / 1e9:.1f} | {tfs:.0f} | {tensor} |")
fl_all = sum(a[3] for a in agg.values())
by_all = sum(a[2] for a in agg.values())
print(f"| **whole pass** | | {len(launches)} | {tot:.1f} | 100% | {by_all / 1e6:.0f} | {by_all / tot / 1e3:.0f} | {by_all / tot / 1e3 / HBM:.2f} | {fl_all / 1e9:.1f} | {fl_all / tot / 1e6:.0f} | {fl_all / tot / 1e6 / TF:.2f} |")
