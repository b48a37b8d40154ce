"""Summarise an ncu launch list (gpu__time_duration.sum + dram bytes per launch) of tools/launch_list.sh.
Usage: python tools/launch_summary.py gpurun_out/launches.csv profiles/rXX_launches_summary.md [profiles/rXX_traffic.json]
Only the LAST forward's launches are kept (everything after the last stem_conv launch minus one forward)."""
import csv, json, re, sys, collections
rows = []
with open(sys.argv[1]) as f:
    rd = csv.reader(l for l in f if l.startswith('"'))
    hdr = next(rd)
    for r in rd:
        d = dict(zip(hdr, r))
        rows.append((int(d["ID"]), d["Kernel Name"], d["Metric Name"], float(d["Metric Value"].replace(",", ""))))
launch = collections.OrderedDict()
for i, name, m, v in rows:
    launch.setdefault(i, {"name": name})[m] = v
ids = list(launch)
stems = [i for i in ids if "stem_conv" in launch[i]["name"]]
first = stems[-1]
sel = [launch[i] for i in ids if i >= first]
def short(n):
    n = re.sub(r"\(.*", "", n).replace("void ", "")
    n = re.sub(r"ftc::_GLOBAL__N__[0-9a-f_]+conv_gemm_t[a-z]+_cu_[0-9a-f]+::", "", n).replace("ftc::", "").replace("(anonymous namespace)::", "")
    return n[-70:]
agg = collections.OrderedDict()
for l in sel:
    a = agg.setdefault(short(l["name"]), [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += l.get("gpu__time_duration.sum", 0.0); a[2] += l.get("dram__bytes_read.sum", 0.0); a[3] += l.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
out = ["# launch list of ONE B=32 bf16 detector forward (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
       "--clock-control none; cold-cache, serialised: compare SHARES, not absolutes)", "",
       f"{len(sel)} launches, {tot/1e6:.2f} ms summed", "",
       "| kernel | launches | total (ms) | share | DRAM read (GB) | DRAM write (GB) |", "|---|---|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {a[0]} | {a[1]/1e6:.3f} | {100*a[1]/tot:.1f} % | {a[2]/1e9:.3f} | {a[3]/1e9:.3f} |")
open(sys.argv[2], "w").write("\n".join(out) + "\n")
print("\n".join(out))
if len(sys.argv) > 3:
    conv = [a for k, a in agg.items() if "conv_gemm_t" in k]
    n = sum(a[0] for a in conv); by = sum(a[2] + a[3] for a in conv); t = sum(a[1] for a in conv)
    json.dump({"kernel": "conv_gemm_tma_kernel + conv_gemm_tc_kernel (tcgen05 implicit-GEMM conv)", "launches_per_forward": n,
               "dram_bytes_per_forward": by, "dram_bytes_per_launch_avg": by / n, "ncu_ms_per_forward": t / 1e6,
               "share_of_forward": t / tot, "batch": 32, "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, tools/launch_list.sh"},
              open(sys.argv[3], "w"), indent=1)
