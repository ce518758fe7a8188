"""Shared-memory instructions with excessive wavefronts (bank conflicts) from an ncu source-page csv.

    ncu -i x.ncu-rep --page source --csv > src.csv ; python tools/ncuconf.py src.csv [kernel section] [N]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sec = rows[starts[k]:(starts[k + 1] if k + 1 < len(starts) else len(rows))]
print(sec[0][1][:120])
h = sec[1]
isrc, iex, iw, ii, ie = (h.index(c) for c in ('Source', 'L1 Wavefronts Shared Excessive', 'L1 Wavefronts Shared',
                                                'L1 Wavefronts Shared Ideal', 'Instructions Executed'))
data = [r for r in sec[2:] if len(r) > max(iex, iw)]
f = lambda x: float(x or 0)
tot_w, tot_x = sum(f(r[iw]) for r in data), sum(f(r[iex]) for r in data)
print(f'shared wavefronts {tot_w:.0f}, excessive {tot_x:.0f} ({100 * tot_x / max(tot_w, 1):.1f} %)')
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
for i in sorted(sorted(range(len(data)), key=lambda i: -f(data[i][iex]))[:n]):
    r = data[i]
    if f(r[iex]) > 0:
        print(f"{i:5d} excess {f(r[iex]):10.0f} of {f(r[iw]):10.0f} (ideal {f(r[ii]):10.0f}) exec {r[ie]:>9s}  {r[isrc].strip()[:80]}")
