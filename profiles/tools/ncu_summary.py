"""Write a short text summary of one kernel from an .ncu-rep (ncu --set full): selected raw metrics + hottest SASS lines.
usage: ncu_summary.py report.ncu-rep "header line" > profiles/<name>.txt"""
import csv, io, subprocess, sys
rep, header = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg"]
print(header)
for k in want:
    for i, name in enumerate(h):
        if name == k or name.endswith("." + k):
            print(f"{name} [{u[i]}] = {v[i]}")
            break
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hh = rows[1]
body = [r for r in rows[2:] if len(r) == len(hh) and r[hh.index("# Samples")].strip().isdigit()]      # (-c > 1 repeats the header)
isamp, isrc = hh.index("# Samples"), hh.index("Source")
stall = [i for i, k in enumerate(hh) if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r[isamp]) for r in body)
st = sorted(((sum(int(r[i]) for r in body), hh[i]) for i in stall), reverse=True)
print("warp stall samples:", ", ".join(f"{n} {100*v/max(tot,1):.1f}%" for v, n in st[:7]))
print("hottest SASS lines (share of samples, dominant stall, instruction):")
for i in sorted(sorted(range(len(body)), key=lambda i: -int(body[i][isamp]))[:8]):
    r = body[i]
    why = max(stall, key=lambda c: int(r[c]))
    print(f"  {100*int(r[isamp])/max(tot,1):5.1f}%  {hh[why]:<18s} {r[isrc].strip()}")
