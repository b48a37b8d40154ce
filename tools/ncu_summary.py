"""Summarise .ncu-rep captures (gpurun_out/) into a markdown table for profiles/.  Usage: ncu_summary.py out.md rep1 rep2 ..."""
import csv, subprocess, sys, io
KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("lts__t_sector_hit_rate.pct", "l2_hit_%")]
out = [f"| capture | kernel | " + " | ".join(k for _, k in KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?").split("(")[0].replace("void ", "")[-60:]
        cells = []
        for k, _ in KEYS:
            v = d.get(k, "")
            try:
                v = f"{float(v):.4g} {u.get(k, '')}".strip()
            except ValueError:
                pass
            cells.append(v)
        out.append(f"| {rep.split('/')[-1]} | `{name}` | " + " | ".join(cells) + " |")
open(sys.argv[1], "w").write("\n".join(out) + "\n")
print("\n".join(out))
