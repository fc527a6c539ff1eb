"""Count the SASS mnemonics that prove Blackwell-native code, per kernel of libv1t_b200.so (no GPU needed):

    python scripts/sass_markers.py > profiles/r2_sass_markers.txt

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (TMA engine, non-tensor form), UTMALDG /
UTMASTG = cp.async.bulk.tensor, SYNCS = mbarrier ops, REDG = red.global, HMMA = legacy mma.sync (must be absent).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "v1t_b200", "libv1t_b200.so")
MARKS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "REDG", "HMMA", "LDGSTS"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for k in MARKS:
            if op.startswith(k):
                counts[cur][k] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"SASS markers per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; sm_100a)")
print(f"{'kernel':<70s} {'instr':>7s} " + " ".join(f"{k:>8s}" for k in MARKS))
rows = []
for (name, c), dn in zip(counts.items(), demangle):
    if not any(c[k] for k in MARKS if k not in ("REDG", "LDGSTS", "SYNCS")) and "--all" not in sys.argv:
        continue
    short = re.sub(r"\(.*", "", dn.replace("(anonymous namespace)::", "").replace("void ", "")).replace("v1t::", "")
    rows.append((short, c))
for short, c in sorted(rows, key=lambda r: -r[1]["UTCHMMA"]):
    print(f"{short[:70]:<70s} {c['_total']:>7d} " + " ".join(f"{c[k]:>8d}" for k in MARKS))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f"{'ALL KERNELS (' + str(len(counts)) + ')':<70s} {tot['_total']:>7d} " + " ".join(f"{tot[k]:>8d}" for k in MARKS))
