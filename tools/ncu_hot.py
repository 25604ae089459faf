"""Summarise an `ncu --page source --csv` SASS dump: hottest instruction ranges, lane utilisation."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot_s = sum(int(r[ix["# Samples"]] or 0) for r in data)
tot_i = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
tot_t = sum(int(r[ix["Thread Instructions Executed"]] or 0) for r in data)
print("instructions", len(data), "samples", tot_s, "warp-inst", tot_i, "avg threads/inst %.2f" % (tot_t / max(tot_i, 1)))
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for r in top:
    s = int(r[ix["# Samples"]] or 0)
    print("%5.2f%% smp  %9s inst  thr %5s  lsb %5s  %s" % (100.0 * s / tot_s, r[ix["Instructions Executed"]], r[ix["Avg. Threads Executed"]],
          r[ix["stall_long_sb"]], r[ix["Source"]][:90]))
