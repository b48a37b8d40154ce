"""Per-kernel counts of the SASS mnemonics that prove which hardware path a kernel uses (B200_PROFILING.md): tcgen05 tensor cores
(UTCHMMA, TMEM loads LDTM / stores STTM), TMA (UTMALDG / UTMASTG tensor tiles, UBLKCP bulk copies), legacy warp-level tensor cores
(HMMA + LDSM), packed fp32 (FFMA2).  Reads the shipped library with cuobjdump; writes a markdown table.

    python tools/sass_summary.py [findtextcenternet_b200/lib/libftc_b200.so] > profiles/r02_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "findtextcenternet_b200", "lib", "libftc_b200.so")
MNEMONICS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDSM", "FFMA2", "MUFU", "REDG", "ATOMG"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = re.sub(r"\(anonymous namespace\)::|ftc::|void ", "", name)
        name = re.sub(r"\(.*", "", name)
        cur = counts.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        cur["_total"] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + "."):
                cur[mn] += 1
print("# SASS summary of libftc_b200.so (sm_100a): which hardware path each kernel is on\n")
print("`cuobjdump -sass` of the shipped library; counts are static instruction counts per kernel (loops not unrolled count once).")
print("UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = TMA bulk copy,")
print("HMMA/LDSM = legacy mma.sync / ldmatrix, FFMA2 = packed fp32 FMA, REDG/ATOMG = global reductions / atomics.\n")
print("| kernel | SASS instr | " + " | ".join(MNEMONICS) + " |")
print("|---|---|" + "---|" * len(MNEMONICS))
rows = sorted(counts.items(), key=lambda kv: (-kv[1]["UTCHMMA"], -kv[1]["UTMALDG"], -kv[1]["HMMA"], kv[0]))
for name, c in rows:
    if not any(c[m] for m in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA")) and c["_total"] < 400:
        continue
    print(f"| `{name[:90]}` | {c['_total']} | " + " | ".join(str(c[m]) if c[m] else "" for m in MNEMONICS) + " |")
tc = [n for n, c in counts.items() if c["UTCHMMA"]]
print(f"\n{len(counts)} kernels in the library; {len(tc)} issue tcgen05.mma (UTCHMMA): " + ", ".join(f"`{n[:60]}`" for n in tc))
