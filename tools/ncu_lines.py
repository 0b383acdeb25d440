#!/usr/bin/env python3
"""Summarise an ncu `--page source --print-source cuda,sass --csv` dump per CUDA source line:
instructions executed, share, average active threads, stall samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0] not in ("", "-"):
        d = dict(zip(hdr[4:], r[4:]))
        out.append((cur_file, r[0], r[1].strip(), d))

def f(d, k):
    try:
        return float(d.get(k, 0) or 0)
    except ValueError:
        return 0.0

tot = sum(f(d, "Instructions Executed") for _, _, _, d in out)
tots = sum(f(d, "# Samples") for _, _, _, d in out)
print(f"total warp instructions {tot:.3e}   stall samples {tots:.0f}")
out.sort(key=lambda x: -f(x[3], "Instructions Executed"))
for file, line, src, d in out[:top]:
    ie = f(d, "Instructions Executed")
    print(f"{100 * ie / tot:5.1f}% inst  {100 * f(d, '# Samples') / max(tots, 1):5.1f}% smp  thr {f(d, 'Avg. Threads Executed'):4.1f}  "
          f"{file}:{line:>4}  {src[:105]}")
