"""Basic-block view of one kernel in an ncu report: runs of SASS instructions with equal execution counts, their share of
the executed warp instructions and of the stall samples, and their opcode mix.
usage: python tools/ncu_blocks.py report.ncu-rep <launch ID> [min share %]"""
import csv
import subprocess
import sys

rep, want = sys.argv[1], int(sys.argv[2])
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
tables, cur = [], None
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        tables.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
# one or two tables per launch (the second repeats the kernel name): take the last table of the wanted launch
groups = []
for tb in tables:
    if groups and groups[-1][-1]["name"] == tb["name"] and len(groups[-1]) < 2:
        groups[-1].append(tb)
    else:
        groups.append([tb])
t = groups[want][-1]
h = t["rows"][0]
ie, src, sa = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
body = t["rows"][1:]
tot = sum(int(r[ie] or 0) for r in body)
ts = sum(int(r[sa] or 0) for r in body)
print(t["name"][:100], "| warp instructions", tot, "| samples", ts)
runs = []
for pos, r in enumerate(body):
    n, s_ = int(r[ie] or 0), int(r[sa] or 0)
    if runs and runs[-1][2] == n:
        runs[-1][1] = pos
        runs[-1][3] += n
        runs[-1][4] += s_
    else:
        runs.append([pos, pos, n, n, s_])
for a, b, n, total, s_ in runs:
    if total > tot * min_share / 100 or s_ > ts * min_share / 100:
        ops = {}
        for r in body[a:b + 1]:
            w = r[src].strip().split()
            o = (w[1] if w[0].startswith("@") else w[0]).split(".")[0]
            ops[o] = ops.get(o, 0) + 1
        mix = " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:9])
        print(f"@{a:4d}-{b:4d} len {b - a + 1:3d} x{n:7d} = {100 * total / tot:5.1f}% instr {100 * s_ / max(ts, 1):5.1f}% samples | {mix}")
