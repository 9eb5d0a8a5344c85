"""Top source lines by warp-stall samples from an `ncu --page source --csv` export: python scripts/src_top.py file.csv [n]"""
import csv, sys
fn = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 28
rows = list(csv.reader(open(fn)))
hdr = next((x for x in rows if "Source" in x), None)
if not hdr:
    print(fn, "no source header", rows[:3]); sys.exit()
i0 = rows.index(hdr)
iS = hdr.index("Source"); iA = hdr.index("Warp Stall Sampling (All Samples)"); iE = hdr.index("Instructions Executed")
data = [x for x in rows[i0 + 1:] if len(x) > iA and x[iA].isdigit()]
tot = sum(int(x[iA]) for x in data)
print("==", fn, "total samples", tot)
for x in sorted(data, key=lambda x: -int(x[iA]))[:n]:
    print(f"{int(x[iA]):7d} {100*int(x[iA])/tot:5.1f}% exec={x[iE]:>10s} {x[iS].strip()[:140]}")
