"""Top stall-sampled SASS lines of one kernel from `ncu -i rep --page source --csv` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Address")
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
i_src, i_samp, i_exec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[i_samp]) for r in data)
print("total samples", tot, "instrs", len(data))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for r in sorted(data, key=lambda r: -int(r[i_samp]))[:n]:
    st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:2]
    print(r[i_samp].rjust(7), f"{100*int(r[i_samp])/max(tot,1):5.1f}%", r[i_exec].rjust(9), r[i_src][:70].ljust(70), st)
