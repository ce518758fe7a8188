"""Markdown table of the largest eager encoder pass in an ncu per-launch csv (tools/gpu_round_profile.sh) and, with
--traffic OUT.json, the measured DRAM bytes per launch of the tcgen05 GEMM family (bench.py's roofline.traffic).

    python tools/launch_summary.py profiles/r01g_launches_eager_b32.csv [--traffic profiles/gemm_traffic.json]
"""
import collections, csv, json, re, sys
path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if r]
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h, data = rows[hi], rows[hi + 1:]
iN, iM, iV, iU, iID = (h.index(c) for c in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
L = collections.OrderedDict()
for r in data:
    d = L.setdefault(r[iID], {'name': r[iN]})
    v, u = float(r[iV].replace(',', '')), r[iU]
    if r[iM].startswith('dram'):
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    elif r[iM].startswith('gpu__time'):
        v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(u, 1)
    d[r[iM]] = v
Ls = list(L.values())
best = None
for s in [i for i, d in enumerate(Ls) if 'patch_embed' in d['name']]:
    ee = [i for i, d in enumerate(Ls) if i > s and 'tokens_nchw' in d['name'] or i > s and 'tokens_to_nchw' in d['name']]
    if ee:
        t = sum(d['gpu__time_duration.sum'] for d in Ls[s:ee[0] + 1])
        if best is None or t > best[0]:
            best = (t, s, ee[0])
_, s, e = best
P = Ls[s:e + 1]
agg = collections.OrderedDict()
for d in P:
    n = re.sub(r'\(.*', '', d['name'])
    for junk in ('void ', 'dlimg::', '(anonymous namespace)::', '<unnamed>::'):
        n = n.replace(junk, '')
    a = agg.setdefault(n, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d['gpu__time_duration.sum']
    a[2] += d['dram__bytes_read.sum']
    a[3] += d['dram__bytes_write.sum']
tot = sum(a[1] for a in agg.values())
print('| kernel | launches | us | share | DRAM read MB | DRAM write MB |\n|---|---:|---:|---:|---:|---:|')
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| `{n}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% | {a[2] / 1e6:.0f} | {a[3] / 1e6:.0f} |')
print(f'| total | {len(P)} | {tot:.1f} | | {sum(a[2] for a in agg.values()) / 1e6:.0f} | {sum(a[3] for a in agg.values()) / 1e6:.0f} |')
if '--traffic' in sys.argv:
    g = [d for d in P if 'gemm_tc_kernel' in d['name'] or 'mlp_fused' in d['name']]
    tb = sum(d['dram__bytes_read.sum'] + d['dram__bytes_write.sum'] for d in g)
    json.dump({"kernel": "gemm_tc_kernel<f16> + mlp_fused_kernel (all %d launches of one eager encoder pass, batch 32)" % len(g),
               "launches": len(g), "dram_bytes_per_launch": tb / len(g), "dram_bytes_total": tb,
               "source": path + ": ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                         "--clock-control none (tools/gpu_round_profile.sh)"},
              open(sys.argv[sys.argv.index('--traffic') + 1], 'w'), indent=1)
