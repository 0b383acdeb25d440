#!/usr/bin/env python3
"""Turn what tools/collect_evidence.sh brought back (gpurun_out/evidence/<tag>_*) into the tracked files under
profiles/: bench JSON lines, per-dataset LZ table, launch list (markdown + csv), one ncu summary per decode
kernel (tools/ncu_summary.py), traffic.json (DRAM bytes per launch, read by bench.py), memcheck logs.
usage: python tools/publish_evidence.py [tag]"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "evidence")
DST = os.path.join(ROOT, "profiles")
DATASET = {"snappy": "tabular_f32", "lz4": "lz4_mixed", "cascaded": "sorted_i64", "bitcomp": "sorted_i64",
           "ans": "lowentropy_bytes"}


def copy(name):
    shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], rows[2] if len(rows) > 2 else rows[1])), dict(zip(rows[0], rows[1]))


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(value) * scale[unit]


def main():
    os.makedirs(DST, exist_ok=True)
    for c in DATASET:
        copy(f"{TAG}_bench_{c}.json")
    copy(f"{TAG}_bench_reference.json")
    copy(f"{TAG}_lz_per_dataset.jsonl")
    copy(f"{TAG}_bench_launches.csv")
    for log in ("memcheck_fuzz", "memcheck_parity"):
        # keep the verdict lines, not the whole pytest transcript
        lines = open(os.path.join(SRC, f"{TAG}_{log}.log")).read().splitlines()
        keep = [ln for ln in lines if "COMPUTE-SANITIZER" in ln or "ERROR SUMMARY" in ln or " passed" in ln or " failed" in ln]
        open(os.path.join(DST, f"{TAG}_{log}.log"), "w").write("\n".join(keep) + "\n")

    # launch list: total time per kernel name
    rows = list(csv.reader(open(os.path.join(SRC, f"{TAG}_bench_launches.csv"))))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hdr_i]
    k_i, v_i = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[hdr_i + 2:]:
        if len(r) > v_i:
            tot[r[k_i]] += float(r[v_i].replace(",", ""))
            cnt[r[k_i]] += 1
    total = sum(tot.values())
    with open(os.path.join(DST, f"{TAG}_bench_launch_list.md"), "w") as f:
        f.write(f"# round {TAG[1:]} — launch list of `python bench.py --steps 3 --warmup 3 --no-cpu --no-extras` (snappy, tabular_f32)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (cold-cache, serialised: compare "
                f"shares, not absolutes). Raw CSV: `{TAG}_bench_launches.csv`.\n\n| kernel | launches | total time | share |\n|---|---|---|---|\n")
        for k, v in tot.most_common():
            f.write(f"| `{k[:80]}` | {cnt[k]} | {v:.1f} ns | {100 * v / total:.1f} % |\n")

    # ncu summaries + DRAM traffic per launch
    traffic = {}
    for c, ds in DATASET.items():
        rep = os.path.join(SRC, f"{TAG}_{c}_{ds}.ncu-rep")
        title = f"round {TAG[1:]} — {c} decompress kernel, 10000 x 64 KB chunks ({ds})"
        md = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, title],
                            capture_output=True, text=True).stdout
        open(os.path.join(DST, f"{TAG}_{c}_{ds}_ncu_summary.md"), "w").write(md)
        val, unit = ncu_raw(rep)
        traffic[f"{c}:{ds}"] = int(to_bytes(val["dram__bytes_read.sum"], unit["dram__bytes_read.sum"]) +
                                   to_bytes(val["dram__bytes_write.sum"], unit["dram__bytes_write.sum"]))
        light = os.path.join(SRC, f"{TAG}_traffic_light_{c}.csv")
        if os.path.exists(light):   # LZ codecs: add the light kernel of the same DecompressAsync
            lrows = list(csv.reader(open(light)))
            hi = next(i for i, r in enumerate(lrows) if r and r[0] == "ID")
            h = lrows[hi]
            for r in lrows[hi + 1:]:
                if len(r) > h.index("Metric Value") and r[h.index("Metric Name")].startswith("dram__bytes"):
                    traffic[f"{c}:{ds}"] += int(to_bytes(r[h.index("Metric Value")].replace(",", ""), r[h.index("Metric Unit")]))
    for name, title in (("snappy_price_walk", "snappy dense block decoder, 10000 x 64 KB chunks (tabular_f32:0, price-walk column)"),
                        ("lz4_runlength_i32", "lz4 light (direct) kernel, 10000 x 64 KB chunks (runlength_i32)"),
                        ("snappy_runlength_i32", "snappy light (direct) kernel, 10000 x 64 KB chunks (runlength_i32)")):
        rep = os.path.join(SRC, f"{TAG}_{name}.ncu-rep")
        if os.path.exists(rep):
            md = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, f"round {TAG[1:]} — {title}"],
                                capture_output=True, text=True).stdout
            open(os.path.join(DST, f"{TAG}_{name}_ncu_summary.md"), "w").write(md)
    traffic["_source"] = (f"profiles/{TAG}_<codec>_<dataset>_ncu_summary.md: dram__bytes_read.sum + dram__bytes_write.sum of one "
                          "`ncu --set full` capture of the named decode kernel on this workload (not measured in the bench run); for snappy / lz4 plus the "
                          f"same two counters of the light kernel of that DecompressAsync ({TAG}_traffic_light_<codec>.csv)")
    json.dump(traffic, open(os.path.join(DST, "traffic.json"), "w"), indent=1)
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
