"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name: launches, summed ms, share.
Usage: python tools/kernel_shares.py launches.csv "title" > summary.md   (profiler times are cold-cache / serialised: shares only)"""
import collections
import csv
import re
import sys

agg = collections.OrderedDict()
with open(sys.argv[1]) as f:
    rd = csv.reader(l for l in f if l.startswith('"'))
    hdr = next(rd)
    for r in rd:
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "")
        name = re.sub(r"ftc::|\(anonymous namespace\)::|_GLOBAL__N__[0-9a-f_]+_cu_[0-9a-f]+::", "", name)[-70:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(d["Metric Value"].replace(",", ""))
tot = sum(a[1] for a in agg.values()) or 1.0
title = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
print(f"# {title} (ncu, cold-cache, serialised: shares only)\n")
print(f"{sum(a[0] for a in agg.values())} launches, {tot / 1e6:.2f} ms summed\n")
print("| kernel | launches | ms | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| `{k}` | {n} | {t / 1e6:.3f} | {100 * t / tot:.1f} % |")
