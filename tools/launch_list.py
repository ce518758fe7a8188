"""Prints an ncu per-launch csv (gpu__time_duration + dram bytes) launch by launch, optionally a slice.
    python tools/launch_list.py file.csv [first [last]]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
L = collections.OrderedDict()
for r in rows[hdr + 1:]:
    d = L.setdefault(int(r[0]), {'name': r[4].split('(')[0].replace('dlimg::', ''), 'grid': r[8], 'block': r[7] if False else r[9] if len(r) > 9 else ''})
    d[r[12]] = float(r[14].replace(',', ''))
a = int(sys.argv[2]) if len(sys.argv) > 2 else 0
b = int(sys.argv[3]) if len(sys.argv) > 3 else max(L) + 1
tot = 0
agg = collections.Counter(); cnt = collections.Counter()
for i, d in L.items():
    if not (a <= i < b):
        continue
    t = d.get('gpu__time_duration.sum', 0) / 1000
    tot += t
    agg[d['name'][:60]] += t; cnt[d['name'][:60]] += 1
    print(f"{i:3d} {d['name'][:60]:60s} {d['grid']:>16s} {t:8.1f} us  rd {d.get('dram__bytes_read.sum', 0) / 1e6:7.1f} wr {d.get('dram__bytes_write.sum', 0) / 1e6:7.1f} MB")
print('total us', round(tot, 1))
for k, v in agg.most_common():
    print(f"  {k:60s} x{cnt[k]:3d} {v:8.1f} us {100 * v / tot:5.1f}%")
