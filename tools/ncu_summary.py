#!/usr/bin/env python3
"""Turn an .ncu-rep (captured with `ncu --set full --clock-control none --import-source on`) into the small
markdown summary committed under profiles/: duration, DRAM bytes, throughput percentages, issue utilisation,
stall mix, and the hottest source lines.   usage: ncu_summary.py report.ncu-rep [title] > profiles/x.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
print(f"# {title}\n")
print(f"source: `{rep.split('/')[-1]}` (ncu --set full --clock-control none --import-source on; cold-cache, serialised replay)\n")
WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "avg active threads / instruction"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (per issue)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
]
units = rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print(f"## kernel `{d.get('Kernel Name', '?')[:90]}`\n")
    print("| metric | value |\n|---|---|")
    for k, name in WANT:
        if k in d:
            print(f"| {name} (`{k}`) | {d[k]} {u.get(k, '')} |")
    try:
        tr = float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])
        print(f"| **DRAM traffic per launch (read+write)** | {tr:.3f} {u.get('dram__bytes_read.sum', '')} |")
    except (KeyError, ValueError):
        pass
    print()
# hottest source lines
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
cur, h, out = None, None, []
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        h = r
    elif h and r[0] not in ("", "-", "Function Name"):
        d = dict(zip(h[4:], r[4:]))
        try:
            out.append((float(d.get("Instructions Executed", 0) or 0), float(d.get("# Samples", 0) or 0), cur, r[0], r[1].strip()))
        except ValueError:
            pass
tot = sum(o[0] for o in out) or 1
tots = sum(o[1] for o in out) or 1
out.sort(reverse=True)
print("## hottest source lines (share of executed warp instructions / of stall samples)\n")
print("| inst % | samples % | line | source |\n|---|---|---|---|")
for ie, sm, f, ln, code in out[:25]:
    code = code.replace("|", "\\|")[:100]
    print(f"| {100 * ie / tot:.1f} | {100 * sm / tots:.1f} | {f}:{ln} | `{code}` |")
