"""Summarises an ncu per-launch csv (tools/gpu_launches.sh): one eager encoder pass, launch by launch.

    python tools/launch_table.py gpurun_out/launches.csv [--step N]
"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
L = collections.OrderedDict()
for r in rows[hdr + 1:]:
    d = L.setdefault(int(r[0]), {'name': r[4].split('(')[0].split('::')[-1], 'grid': r[8]})
    d[r[12]] = float(r[14].replace(',', ''))
starts = [i for i in L if 'conv1_preprocess' in L[i]['name'] or 'patch_embed' in L[i]['name']]
step = int(sys.argv[sys.argv.index('--step') + 1]) if '--step' in sys.argv else 1
s = starts[step]
e = starts[step + 1] if step + 1 < len(starts) else max(L) + 1
tot = 0
agg = collections.Counter()
for i in range(s, e):
    d = L[i]
    t = d.get('gpu__time_duration.sum', 0) / 1000
    tot += t
    agg[d['name'][:40]] += t
    print(f"{i - s:3d} {d['name'][:44]:44s} {d['grid']:>16s} {t:8.1f} us  rd {d.get('dram__bytes_read.sum', 0) / 1e6:7.1f} wr {d.get('dram__bytes_write.sum', 0) / 1e6:7.1f} MB")
print('total us', round(tot, 1))
for k, v in agg.most_common():
    print(f"  {k:40s} {v:8.1f} us {100 * v / tot:5.1f}%")
