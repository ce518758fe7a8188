"""Top SASS instructions by stall samples / executed count from an ncu source-page csv.

    ncu -i x.ncu-rep --page source --csv > src.csv
    python tools/ncusrc.py src.csv <kernel section index> [N]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sec = rows[starts[k]:(starts[k + 1] if k + 1 < len(starts) else len(rows))]
print(sec[0][1][:120])
h = sec[1]
isrc, isamp, iexec = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
data = [r for r in sec[2:] if len(r) > max(isamp, iexec)]
tot_s = sum(int(r[isamp]) for r in data)
tot_e = sum(int(r[iexec]) for r in data)
print('sections', len(starts), 'total samples', tot_s, 'total warp-instr executed', tot_e, 'static instrs', len(data))
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:n]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[c]), h[c][6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[isamp]):6d} {100*int(r[isamp])/tot_s:5.1f}%  exec {int(r[iexec]):9d}  {r[isrc].strip()[:70]:70s} {st}")
