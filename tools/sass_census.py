"""Per-kernel census of the SASS mnemonics that prove the Blackwell-native path (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG / UBLKCP
(TMA loads / stores / reductions / 1-D bulk copies), HMMA (legacy mma.sync — expected absent), plus registers per thread.
Runs without a GPU:  python tools/sass_census.py > profiles/<tag>_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "graph-gpt_b200", "libggpt_b200.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UBLKRED", "SYNCS", "LDGSTS", "MUFU.EX2",
         "HMMA", "STG.E.ENL2.256", "RED.E", "REDG"]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = {}
kern, counts, order = None, collections.defaultdict(collections.Counter), []
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        order.append(kern)
        continue
    if kern is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[kern]["_total"] += 1
        for w in WATCH:
            if op.startswith(w):
                counts[kern][w] += 1
names = subprocess.run(["cu++filt"] + order, capture_output=True, text=True).stdout.splitlines() if order else []
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+)", line)
    if m and cur:
        regs[cur] = int(m.group(1))
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)} (sm_100a): instruction counts per kernel; columns are SASS mnemonic prefixes")
cols = [w for w in WATCH if any(counts[k][w] for k in order)] + [w for w in ("HMMA",) if not any(counts[k][w] for k in order)]
print("kernel".ljust(58) + "regs".rjust(5) + "instr".rjust(7) + "".join(c.rjust(max(9, len(c) + 1)) for c in cols))
for k, n in sorted(zip(order, names), key=lambda kn: kn[1]):
    short = n.replace("(int)", "").replace("(bool)", "").replace("void ", "").replace("ggpt::", "")
    short = short[:short.rfind("(")] if "(" in short else short
    if not any(counts[k][w] for w in WATCH if w not in ("LDGSTS", "RED.E", "REDG", "MUFU.EX2")) and "--all" not in sys.argv:
        continue
    print(short[:57].ljust(58) + str(regs.get(k, "")).rjust(5) + str(counts[k]["_total"]).rjust(7)
          + "".join(str(counts[k][c] or ".").rjust(max(9, len(c) + 1)) for c in cols))
tot = collections.Counter()
for k in order:
    tot.update(counts[k])
print("\n# whole library: " + ", ".join(f"{w} {tot[w]}" for w in WATCH))
print(f"# kernels in the library: {len(order)}; listed above: those that use tcgen05 / TMEM / TMA / mbarrier instructions (--all lists every kernel)")
