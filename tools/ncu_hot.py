"""Hottest SASS instructions of one kernel in an ncu report by warp-stall samples (needs --set full --import-source on).
usage: python tools/ncu_hot.py report.ncu-rep <launch ID in the report> [top n] [src]   (src: CUDA source lines instead of SASS)"""
import csv
import subprocess
import sys

rep, want = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
tables, cur = [], None
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        tables.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
groups = []
for tb in tables:  # one or two tables per launch (the second repeats the kernel name)
    if groups and groups[-1][-1]["name"] == tb["name"] and len(groups[-1]) < 2:
        groups[-1].append(tb)
    else:
        groups.append([tb])
t = groups[want][-1 if "src" in sys.argv[3:] else 0]
h = t["rows"][0]
si, src = h.index("# Samples"), h.index("Source")
body = [r for r in t["rows"][1:] if len(r) > si]
tot = sum(int(r[si] or 0) for r in body)
print(t["name"][:110], "| total samples", tot, "| instructions", len(body))
for pos, r in enumerate(body):
    r.append(pos)  # position in program order
for r in sorted(body, key=lambda r: -int(r[si] or 0))[:top]:
    print(f"{int(r[si]):6d} {100.0 * int(r[si]) / max(tot, 1):5.1f}%  @{r[-1]:5d}  {r[src].strip()[:110]}")
