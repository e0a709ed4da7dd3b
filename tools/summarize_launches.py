"""Summarises an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list: per kernel name the launch
count, total / average duration, share of the total kernel time and (when the DRAM counters were collected) the
DRAM bytes per launch.
usage: python tools/summarize_launches.py launches.csv [frames] [--last-frames N --marker KERNEL]
  --last-frames N --marker K : keep only the launches from N occurrences of kernel K before the end, backed up to
                               the preceding voxel_hash_kernel (the start of that frame's preprocessing)"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    frames = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 1.0
    last = int(sys.argv[sys.argv.index("--last-frames") + 1]) if "--last-frames" in sys.argv else 0
    marker = sys.argv[sys.argv.index("--marker") + 1] if "--marker" in sys.argv else "frame_begin_kernel"
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    if last:
        ids = []
        for r in rows:
            if r["ID"] not in ids:
                ids.append(r["ID"])
        first_row = {i: next(r for r in rows if r["ID"] == i) for i in ids}
        marks = [i for i in ids if marker in first_row[i]["Kernel Name"]]
        start = ids.index(marks[-last])
        while start > 0 and "voxel_hash_kernel" not in first_row[ids[start]]["Kernel Name"]:
            start -= 1
        keep = set(ids[start:])
        rows = [r for r in rows if r["ID"] in keep]
        frames = float(last)
    per = defaultdict(lambda: defaultdict(float))
    count = defaultdict(int)
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("pls::<unnamed>::", "").strip()
        metric, val = r["Metric Name"], float(r["Metric Value"].replace(",", "") or 0)
        unit = r.get("Metric Unit", "")
        if metric == "gpu__time_duration.sum":
            val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
            count[name] += 1
        elif metric.startswith("dram__bytes"):
            val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per[name][metric] += val
    total = sum(v["gpu__time_duration.sum"] for v in per.values())
    print(f"frames captured: {frames}; total kernel time per frame {total / frames:.1f} us "
          f"(ncu, cold-cache, serialised: compare SHARES, not absolutes)")
    for name, v in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        us = v["gpu__time_duration.sum"]
        n = max(count[name], 1)
        line = (f"{name[:60]:<60} n/frame={n / frames:6.1f} us/frame={us / frames:9.1f} avg_us={us / n:8.2f} "
                f"share={100 * us / total:5.1f}%")
        dram = v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0)
        if dram:
            line += f" dram_MB/launch={dram / n / 1e6:8.3f} dram_GB/s={dram / n / (us / n * 1e-6) / 1e9:8.1f}"
        print(line)


if __name__ == "__main__":
    main()
