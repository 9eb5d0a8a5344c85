"""Summarise an .ncu-rep: key raw metrics + stall samples by source line (needs -lineinfo)."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread ",
        "launch__occupancy_limit", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum ", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg ", "launch__grid_size", "launch__waves",
        "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_lsu", "sm__pipe_fma_cycles_active.avg.pct", "sm__pipe_alu_cycles_active.avg.pct", "smsp__issue_active.avg.pct"]
for h, u, v in zip(r[0], r[1], r[2]):
    hh = h + " "
    if any(k in hh for k in keys) and ".max" not in h and ".min" not in h and ".sum.pct" not in h:
        print(f"{h:90s} {u:12s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next((x for x in rows if "Source" in x), None)
if hdr:
    i0 = rows.index(hdr)
    iS, iA = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    iE = hdr.index("Instructions Executed")
    data = [x for x in rows[i0 + 1:] if len(x) > iA and x[iA].isdigit()]
    tot = sum(int(x[iA]) for x in data)
    print("\ntotal samples", tot)
    for x in sorted(data, key=lambda x: -int(x[iA]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
        print(f"{int(x[iA]):7d} {100*int(x[iA])/tot:5.1f}%  exec={x[iE]:>10s}  {x[iS].strip()[:110]}")
