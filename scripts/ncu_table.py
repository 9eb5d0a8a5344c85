"""Markdown table from an `ncu --page raw --csv` export: per kernel launch duration, DRAM traffic and achieved GB/s (and its share of the
measured HBM peak), tensor-pipe utilisation, issue-active, resident warps, registers, top stall reasons.
    python scripts/ncu_table.py raw.csv[.gz] [title] > profiles/xxx.md"""
import csv, gzip, io, json, os, sys
path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
f = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)
r = [row for row in csv.reader(f) if row and not row[0].startswith("==")]
hdr, units, rows = r[0], r[1], r[2:]
peaks = {"hbm": 6485.2, "tf": 1422.7}
pp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pp):
    d = json.load(open(pp)); peaks = {"hbm": d["hbm_gbs"], "tf": d.get("bf16_tflops_sustained", d["bf16_tflops"])}
def col(n):
    return hdr.index(n) if n in hdr else None
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return float("nan")
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
c = {k: col(v) for k, v in dict(name="Kernel Name", t="gpu__time_duration.sum", dr="dram__bytes_read.sum", dw="dram__bytes_write.sum",
     tensor="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", issue="smsp__issue_active.avg.pct_of_peak_sustained_active",
     warps="sm__warps_active.avg.pct_of_peak_sustained_active", regs="launch__registers_per_thread", grid="launch__grid_size",
     l2hit="lts__t_sector_hit_rate.pct").items()}
stall = [(i, h.split("issue_stalled_")[1].split("_per")[0]) for i, h in enumerate(hdr)
         if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
print(f"# {title}\n")
print(f"`ncu --set full --clock-control none` on one B200; DRAM GB/s = (dram read + write bytes) / duration, against the measured copy bandwidth "
      f"{peaks['hbm']:.0f} GB/s (MEASURED_PEAKS.json).  Cold-cache, serialised launches: for shares and pipe utilisation, not for bench values.\n")
print("| # | kernel | grid | regs | us | DRAM MB | DRAM GB/s | % of HBM peak | tensor pipe % | issue active % | warps active % | top stalls (cycles / issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for n, row in enumerate(rows):
    name = row[c["name"]].replace("es3::", "").replace("(anonymous namespace)::", "").replace("void ", "")
    name = name.split("(")[0][:58]
    t = num(row[c["t"]]) * mult.get(units[c["t"]], 1.0)
    by = num(row[c["dr"]]) * mult.get(units[c["dr"]], 1) + num(row[c["dw"]]) * mult.get(units[c["dw"]], 1)
    gbs = by / 1e9 / (t * 1e-6) if t > 0 else float("nan")
    st = sorted(((num(row[i]), k) for i, k in stall), reverse=True)[:3]
    print(f"| {n} | `{name}` | {row[c['grid']]} | {row[c['regs']]} | {t:.1f} | {by / 1e6:.1f} | {gbs:.0f} | {100 * gbs / peaks['hbm']:.1f} | "
          f"{num(row[c['tensor']]):.1f} | {num(row[c['issue']]):.1f} | {num(row[c['warps']]):.1f} | " + ", ".join(f"{k} {v:.1f}" for v, k in st) + " |")
