"""Per-kernel counts of the SASS mnemonics that prove which hardware paths a kernel uses (B200_PROFILING.md):
UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA load / store), LDTM / STTM (TMEM load / store), UTCBAR (tcgen05.commit),
HMMA (mma.sync), MOVM (movmatrix), LDSM (ldmatrix), UCGABAR (cluster barrier), plus the instruction total.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections, os, re, subprocess, sys
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dlimgedit_b200", "libdlimgedit.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
keys = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "HMMA", "MOVM", "LDSM", "UCGABAR", "LDGSTS", "MUFU"]
cur, rows, idx = None, collections.OrderedDict(), 0
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = names[idx] if idx < len(names) else m.group(1)
        idx += 1
        rows[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        rows[cur]["total"] += 1
        for k in keys:
            if m.group(1).startswith(k):
                rows[cur][k] += 1
print(f"# {os.path.basename(LIB)}: SASS mnemonic counts per kernel (sm_100a)")
print("kernel".ljust(78) + "total".rjust(7) + "".join(k.rjust(9) for k in keys))
tot = collections.Counter()
for name, c in rows.items():
    short = name.replace("(anonymous namespace)::", "").replace("dlimg::", "").replace("void ", "")
    short = re.sub(r"\(.*", "", short)
    print(short[:77].ljust(78) + str(c["total"]).rjust(7) + "".join((str(c[k]) if c[k] else ".").rjust(9) for k in keys))
    tot.update(c)
print("ALL".ljust(78) + str(tot["total"]).rjust(7) + "".join(str(tot[k]).rjust(9) for k in keys))
