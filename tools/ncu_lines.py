"""Per-source-line totals from `ncu --page source --csv --print-source cuda,sass`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
cur_file = None
hdr = None
out = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or r[0] in ("Function Name",) or r[0] == "":
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    try:
        inst = int(r[hdr["Instructions Executed"]] or 0)
        thr = int(r[hdr["Thread Instructions Executed"]] or 0)
        smp = int(r[hdr["# Samples"]] or 0)
    except (ValueError, IndexError):
        continue
    out.append((cur_file, line, r[1].strip()[:80], inst, thr, smp))
ti = sum(o[3] for o in out); ts = sum(o[5] for o in out)
print("total warp-inst %.3e samples %d" % (ti, ts))
for o in sorted(out, key=lambda o: -o[3])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%5.2f%% inst %5.2f%% smp  thr %4.1f  %s:%d  %s" % (100.0 * o[3] / ti, 100.0 * o[5] / ts, o[4] / max(o[3], 1), o[0], o[1], o[2]))
