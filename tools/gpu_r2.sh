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
    timeout 400 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
    tail -c 1500 gpurun_out/${TAG}_bench_default.json; tail -3 gpurun_out/${TAG}_bench_default.err ;;
launches)
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_cfg2.csv \
        python bench.py --quick --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
    python tools/summarize_launches.py gpurun_out/${TAG}_launches_cfg2.csv 1 --last-frames 20 > gpurun_out/${TAG}_launches_cfg2_summary.txt 2>&1
    head -40 gpurun_out/${TAG}_launches_cfg2_summary.txt ;;
ncukd)
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:'kd_nn_warp_kernel|kd_normals_warp_kernel|kd_residual_kernel' \
        --launch-skip ${SKIP:-240} --launch-count ${COUNT:-9} -f -o gpurun_out/${TAG}_kd python bench.py --quick --steps 3 --warmup 24 > gpurun_out/${TAG}_ncu_kd.log 2>&1
    tail -2 gpurun_out/${TAG}_ncu_kd.log
    ncu -i gpurun_out/${TAG}_kd.ncu-rep --page raw --csv > gpurun_out/${TAG}_kd_raw.csv 2>/dev/null
    ls -la gpurun_out/${TAG}_kd* ;;
kdtimes)
    timeout 200 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,smsp__inst_executed.sum --clock-control none -k regex:'kd_|frame_begin' \
        --launch-skip ${SKIP:-400} --launch-count ${COUNT:-60} --csv --log-file gpurun_out/${TAG}_kdtimes.csv \
        python bench.py --quick --steps 4 --warmup 24 > gpurun_out/${TAG}_kdtimes.log 2>&1
    python - <<PY
import csv
lines=[l for l in open("gpurun_out/${TAG}_kdtimes.csv") if l.startswith('"')]
rows=list(csv.DictReader(lines))
by={}
for r in rows:
    by.setdefault(r["ID"],{"name":r["Kernel Name"].split("(")[0].split("::")[-1]})[r["Metric Name"]]=r["Metric Value"]
for i,v in list(by.items())[:60]:
    print(v["name"][:28].ljust(28), "us", float(v.get("gpu__time_duration.sum","0").replace(",",""))/1e3, "cyc", v.get("sm__cycles_active.avg"), "inst", v.get("smsp__inst_executed.sum"))
PY
    ;;
kdprof)
    timeout 200 python tools/kd_profile.py 2>&1 | tail -9 | tee -a gpurun_out/${TAG}_kdprof.log ;;
cellsweep)
    for cell in ${CELLS:-0.16 0.2 0.25 0.3}; do
        echo "== PLS_KD_CELL=$cell"
        PLS_KD_CELL=$cell timeout 200 python tools/kd_profile.py 0 7 9 2>&1 | tail -3 | tee -a gpurun_out/${TAG}_cellsweep.log
    done ;;
projtests)
    timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 200 -rf -p no:cacheprovider -k "proj or cfg3 or cfg5 or a5 or a6 or neighbors" > gpurun_out/${TAG}_pytest_proj.log 2>&1
    tail -8 gpurun_out/${TAG}_pytest_proj.log ;;
projsweep)
    for kd in ${KDIRECTS:-4 8}; do for stg in ${STAGES:-2 3}; do
        echo "== PLS_PROJ_KDIRECT=$kd PLS_PROJ_STAGES=$stg"
        PLS_PROJ_KDIRECT=$kd PLS_PROJ_STAGES=$stg timeout 200 python tools/profile_proj.py 128 4096 26 20 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_projsweep.log
    done; done ;;
traffic)
    M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
    timeout 600 ncu --replay-mode application --cache-control none --clock-control none --metrics $M \
        -k regex:'kd_nn_verify_kernel|kd_nn_warp_kernel|kd_normals_warp_kernel|kd_residual_kernel' --launch-skip ${KDSKIP:-330} --launch-count ${KDCOUNT:-90} \
        --csv --log-file gpurun_out/${TAG}_traffic_kd.csv python bench.py --quick --steps 8 --warmup 24 > gpurun_out/${TAG}_traffic_kd.log 2>&1
    timeout 600 ncu --replay-mode application --cache-control none --clock-control none --metrics $M -k regex:'proj_icp_tma_kernel' \
        --launch-skip 400 --launch-count 40 --csv --log-file gpurun_out/${TAG}_traffic_cfg5.csv python tools/profile_proj.py 128 4096 26 20 > gpurun_out/${TAG}_traffic_cfg5.log 2>&1
    timeout 600 ncu --replay-mode application --cache-control none --clock-control none --metrics $M -k regex:'proj_icp_tma_kernel' \
        --launch-skip 150 --launch-count 60 --csv --log-file gpurun_out/${TAG}_traffic_cfg3.csv python tools/profile_proj.py 128 2048 26 10 > gpurun_out/${TAG}_traffic_cfg3.log 2>&1
    python tools/traffic_summary.py ${TAG} ${COMMIT:-unknown} ;;
ncuresident)
    for mb in ${RESIDENT:-0 64}; do
        PLS_PROJ_RESIDENT_MB=$mb timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum --clock-control none --cache-control none -k regex:'proj_icp_tma_kernel' \
            --launch-skip ${SKIP:-400} --launch-count ${COUNT:-4} --csv --log-file gpurun_out/${TAG}_resident_${mb}.csv python tools/profile_proj.py 128 4096 26 20 > gpurun_out/${TAG}_ncu_resident.log 2>&1
        python - <<PY
import csv
lines=[l for l in open("gpurun_out/${TAG}_resident_${mb}.csv") if l.startswith('"')]
by={}
for r in csv.DictReader(lines):
    by.setdefault(r["ID"],{})[r["Metric Name"]]=r["Metric Value"]
for i,v in by.items():
    print("resident $mb MB:", {k.split("__")[-1][:40]: x for k,x in v.items()})
PY
    done ;;
residentsweep)
    for mb in ${RESIDENT:-0 32 48 56 64 80 96}; do
        echo "== PLS_PROJ_RESIDENT_MB=$mb"
        PLS_PROJ_RESIDENT_MB=$mb timeout 200 python tools/profile_proj.py 128 4096 26 20 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_residentsweep.log
    done
    for mb in ${RESIDENT3:-0 32 64}; do
        echo "== cfg3 PLS_PROJ_RESIDENT_MB=$mb"
        PLS_PROJ_RESIDENT_MB=$mb timeout 200 python tools/profile_proj.py 128 2048 26 10 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_residentsweep.log
    done ;;
ncuproj)
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:'proj_icp_tma_kernel' \
        --launch-skip ${SKIP:-400} --launch-count ${COUNT:-2} -f -o gpurun_out/${TAG}_proj python tools/profile_proj.py 128 4096 26 20 > gpurun_out/${TAG}_ncu_proj.log 2>&1
    tail -2 gpurun_out/${TAG}_ncu_proj.log
    ncu -i gpurun_out/${TAG}_proj.ncu-rep --page raw --csv > gpurun_out/${TAG}_proj_raw.csv 2>/dev/null
    ncu -i gpurun_out/${TAG}_proj.ncu-rep --page source --csv > gpurun_out/${TAG}_proj_source.csv 2>/dev/null
    ls -la gpurun_out/${TAG}_proj* ;;
newtests)
    timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 200 -rf -p no:cacheprovider -k "a3 or a9 or nan or projective_small or cfg5" > gpurun_out/${TAG}_pytest_new.log 2>&1
    tail -12 gpurun_out/${TAG}_pytest_new.log ;;
mgpu)
    NG=$(python -c "import torch; print(torch.cuda.device_count())")
    echo "== multi-GPU on $NG GPUs"
    timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu --timeout 400 -rf -s -p no:cacheprovider > gpurun_out/${TAG}_pytest_mgpu.log 2>&1
    grep -E "mgpu_check|passed|failed|error" gpurun_out/${TAG}_pytest_mgpu.log | tail -12 ;;
mcheck)
    NG=$(python -c "import torch; print(torch.cuda.device_count())")
    for spec in ${CHECKS:-"kdtree p2p stop" "projective p2p fixed"}; do
        timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29547 \
            tools/mgpu_check.py $spec 2>&1 | grep -E "mgpu_check|Error|error" | tail -3 | tee -a gpurun_out/${TAG}_mcheck_n${NG}.log
    done ;;
mbench)
    NG=$(python -c "import torch; print(torch.cuda.device_count())")
    for n in ${NS:-2}; do
        [ "$n" -le "$NG" ] || continue
        for comm in ${COMMS:-p2p}; do
            timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 \
                bench.py --gpus $n --steps 20 --warmup 5 --comm $comm ${BENCHARGS:-} > gpurun_out/${TAG}_bench_n${n}_${comm}.json 2> gpurun_out/${TAG}_bench_n${n}_${comm}.err
            python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_n${n}_${comm}.json").read().strip().splitlines()[-1])
    print("N=$n $comm value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "parity", d["config"].get("sharded_vs_single"))
    for k, v in d["config"].get("extra_workloads", {}).items():
        print("  ", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a in ("ms_per_frame", "ms_per_registration", "kernel_avg_us", "frac_of_measured_hbm_peak", "tiles_sharded", "queries_sharded", "exchange")})
except Exception as e:
    print("bench N=$n $comm failed", e)
PY
            tail -3 gpurun_out/${TAG}_bench_n${n}_${comm}.err
        done
    done ;;
prioab)
    for rep in 1 2; do for pr in 1 0; do
        echo "== PLS_STREAM_PRIORITY=$pr"
        PLS_STREAM_PRIORITY=$pr timeout 120 python tools/e2e_breakdown.py 2>&1 | tail -8 | tee -a gpurun_out/${TAG}_prioab.log
        PLS_STREAM_PRIORITY=$pr timeout 120 python bench.py --quick --steps 40 --warmup 24 2>/dev/null | tail -1 | cut -c1-200
    done; done ;;
e2ebreak)
    timeout 120 python tools/e2e_breakdown.py 2>&1 | tail -8 | tee gpurun_out/${TAG}_e2e_breakdown.log ;;
quicktime)
    timeout 120 python tools/quick_time.py 40 tensor 2>&1 | tail -4 | tee gpurun_out/${TAG}_quicktime.log ;;
stats)
    PLS_KD_STATS=1 timeout 120 python tools/quick_time.py 30 tensor 2>&1 | tail -6 | tee gpurun_out/${TAG}_kd_stats.log ;;
quick)
    for rep in 1 2; do
        timeout 120 python bench.py --quick --steps 40 --warmup 24 2>/dev/null | tail -1 | tee -a gpurun_out/${TAG}_quick.log
    done ;;
esac
done
