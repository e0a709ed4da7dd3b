#!/bin/bash
# Round-2 GPU session helper (outputs under gpurun_out/).  Stages by name, e.g.
#   bash tools/gpu_r2.sh tests bench launches
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${TAG:-r2}
for stage in "$@"; do
case $stage in
tests)
    timeout 600 python -m pytest tests -q -m gpu --timeout 240 -rf -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1
    tail -25 gpurun_out/${TAG}_pytest_gpu.log ;;
kdtests)
    timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 200 -rf -p no:cacheprovider -k "kd or icp or cfg2 or cfg4 or nan or reinit or empty" > gpurun_out/${TAG}_pytest_kd.log 2>&1
    tail -25 gpurun_out/${TAG}_pytest_kd.log ;;
bench)
    timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_s20w5.json 2> gpurun_out/${TAG}_bench.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_s20w5.json"))
    print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 4),
          "launches", d["gpu_launches"], "kernels", {k: round(v, 4) for k, v in d["kernels"].items()}, "roofline us", round(d["roofline"]["avg_us"], 2),
          "iters", d["config"].get("iters_mean"))
except Exception as e:
    print("bench failed", e)
PY
    tail -3 gpurun_out/${TAG}_bench.err ;;
benchlong)
    timeout 300 python bench.py --no-cpu > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
    tail -c 1500 gpurun_out/${TAG}_bench_default.json; tail -3 gpurun_out/${TAG}_bench_default.err ;;
launches)
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_cfg2.csv \
        python bench.py --quick --steps 3 --warmup 24 > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
    python tools/summarize_launches.py gpurun_out/${TAG}_launches_cfg2.csv 3 > gpurun_out/${TAG}_launches_cfg2_summary.txt 2>&1
    head -40 gpurun_out/${TAG}_launches_cfg2_summary.txt ;;
quick)
    for rep in 1 2; do
        timeout 120 python bench.py --quick --steps 40 --warmup 24 2>/dev/null | tail -1 | tee -a gpurun_out/${TAG}_quick.log
    done ;;
esac
done
