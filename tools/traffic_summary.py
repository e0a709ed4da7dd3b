"""Aggregates the ncu application-replay captures of `tools/gpu_r2.sh traffic` into the JSON bench.py reads
(profiles/r2_traffic.json): DRAM bytes (read + write) per EXECUTED launch of each kernel family, with L2 in its
natural in-pipeline state (--cache-control none, --replay-mode application: no save/restore between passes).
    python tools/traffic_summary.py <tag> <commit>"""
import csv, json, sys

tag, commit = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "unknown"


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    by = {}
    for r in csv.DictReader(lines):
        e = by.setdefault(int(r["ID"]), {"name": r["Kernel Name"].split("(")[0].split("::")[-1]})
        e[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        e[r["Metric Name"] + ".unit"] = r["Metric Unit"]
    out = []
    for _, e in sorted(by.items()):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = e.get("dram__bytes_read.sum", 0.0) * scale.get(e.get("dram__bytes_read.sum.unit", "byte"), 1.0)
        wr = e.get("dram__bytes_write.sum", 0.0) * scale.get(e.get("dram__bytes_write.sum.unit", "byte"), 1.0)
        us = e.get("gpu__time_duration.sum", 0.0) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(e.get("gpu__time_duration.sum.unit", "ns"), 1e-3)
        out.append((e["name"], rd + wr, us))
    return out


res = {"commit": commit, "method": "ncu --replay-mode application --cache-control none --clock-control none; "
                                     "dram__bytes_read.sum + dram__bytes_write.sum per executed launch"}
try:
    kd = launches(f"gpurun_out/{tag}_traffic_kd.csv")
    executed = [k for k in kd if k[0] == "kd_residual_kernel" and k[2] > 6.0]   # a no-op launch returns in ~2 us
    total = sum(b for _, b, _ in kd)
    res["kd_iteration_bytes_per_launch"] = total / max(len(executed), 1)
    res["kd_executed_iterations_captured"] = len(executed)
    res["kd_by_kernel_bytes_per_iteration"] = {n: sum(b for m, b, _ in kd if m == n) / max(len(executed), 1)
                                               for n in sorted({m for m, _, _ in kd})}
except Exception as e:
    res["kd_error"] = str(e)
for name in ("cfg3", "cfg5"):
    try:
        pj = [k for k in launches(f"gpurun_out/{tag}_traffic_{name}.csv") if k[2] > 8.0]
        res[f"{name}_proj_bytes_per_launch"] = sum(b for _, b, _ in pj) / max(len(pj), 1)
        res[f"{name}_proj_launches_captured"] = len(pj)
        res[f"{name}_proj_avg_us_under_ncu"] = sum(u for _, _, u in pj) / max(len(pj), 1)
    except Exception as e:
        res[f"{name}_error"] = str(e)
json.dump(res, open(f"gpurun_out/{tag}_traffic.json", "w"), indent=1)
print(json.dumps(res, indent=1))
