"""Summarise `ncu -i X.ncu-rep --page source --csv` output: total samples, stall-reason totals, hottest SASS lines.
usage: ncu_top.py file.csv [top=25]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = rows[1]
body = [r for r in rows[2:] if len(r) == len(h)]
isamp, isrc, iex = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
stall = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r[isamp]) for r in body)
print("kernel:", rows[0][1], " SASS lines:", len(body), " samples:", tot, " warp-instr executed:", sum(int(r[iex]) for r in body))
st = sorted(((sum(int(r[i]) for r in body), h[i]) for i in stall), reverse=True)
print("stalls:", ", ".join(f"{n} {100*v/max(tot,1):.1f}%" for v, n in st[:8]))
idx = sorted(range(len(body)), key=lambda i: -int(body[i][isamp]))[:top]
for i in sorted(idx):
    r = body[i]
    why = max(stall, key=lambda c: int(r[c]))
    print(f"{i:6d} {int(r[isamp]):7d} ({100*int(r[isamp])/max(tot,1):4.1f}%) ex={r[iex]:>9s} {h[why]:<18s} {r[isrc].strip()}")
