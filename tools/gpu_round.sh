#!/bin/bash
# ROUND-1 evidence script, kept so that profiles/r1_* stay reproducible at their commits (tools/gpu_r2.sh is the current
# one).  The A/B stages `ab`, `later` and `abkeys` set toggles of kernels that were replaced in round 2 (PLS_KD_NGROUP*):
# on the current tree they run the default build twice.
# One GPU-box session of the round (outputs under gpurun_out/).  Stages are picked by name:
#   bash tools/gpu_round.sh next subset launches nextprof ncufull full bench
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for stage in "$@"; do
case $stage in
next)
    echo "== new-row parity tests"
    timeout 300 python -m pytest tests/test_next_rows_gpu.py -q -m gpu --timeout 150 -rf > gpurun_out/pytest_next_rows.log 2>&1
    tail -25 gpurun_out/pytest_next_rows.log ;;
subset)
    echo "== parity tests touching the files changed this session (gn.cu, grid_sample.cu)"
    timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 120 -rf -k "gn_ or align or a1_ or a16 or preprocessing or small_vs_reference" \
        > gpurun_out/pytest_subset.log 2>&1
    tail -8 gpurun_out/pytest_subset.log ;;
full)
    echo "== full gpu suite"
    timeout 540 python -m pytest tests -q -m gpu --timeout 240 -rf > gpurun_out/pytest_gpu.log 2>&1
    tail -15 gpurun_out/pytest_gpu.log ;;
launches)
    echo "== launch list cfg2 (3 steady-state frames)"
    timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg2.csv \
        python bench.py --quick --steps 3 --warmup 24 > gpurun_out/bench_under_ncu.log 2>&1
    tail -2 gpurun_out/bench_under_ncu.log; wc -l gpurun_out/launches_cfg2.csv ;;
nextprof)
    echo "== next-row kernels: durations + DRAM bytes"
    timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/launches_next_rows.csv python tools/profile_next_rows.py > gpurun_out/next_rows_profile.json 2> gpurun_out/next_rows_profile.err
    tail -3 gpurun_out/next_rows_profile.json; tail -3 gpurun_out/next_rows_profile.err ;;
ncufull)
    echo "== ncu --set full: kd correspondence kernels of two steady-state frames"
    timeout 300 ncu --set full --clock-control none --import-source on \
        -k regex:'kd_nn_group_kernel|kd_normals_group_kernel|kd_residual_kernel' --launch-skip 264 --launch-count 24 -f -o gpurun_out/r1_kd_group \
        python bench.py --quick --steps 3 --warmup 24 > gpurun_out/ncu_full.log 2>&1
    tail -3 gpurun_out/ncu_full.log
    ncu -i gpurun_out/r1_kd_group.ncu-rep --page raw --csv > gpurun_out/r1_kd_group_raw.csv 2>/dev/null
    ls -la gpurun_out/r1_kd_group* ;;
final)
    echo "== full gpu suite, new tests first"
    timeout 450 python -m pytest tests/test_next_rows_gpu.py tests/test_gpu_parity.py tests/test_multi_gpu.py -q -m gpu --timeout 200 -rf \
        --durations=8 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
    tail -30 gpurun_out/pytest_gpu.log ;;
ab)
    echo "== A/B: lanes per pending normal (device-resident frames only)"
    for ng in 4 8; do
        PLS_KD_NGROUP=$ng timeout 90 python bench.py --quick --steps 40 --warmup 24 2>/dev/null | tail -1 | \
            python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ngroup $ng ms/frame', round(d['ms_per_step'],4), 'iters', d['iters_mean'])" \
            | tee -a gpurun_out/ab_ngroup.log
    done ;;
abkeys)
    echo "== A/B: 40-bit compact sort keys in the grid sample vs raw 64-bit keys (device-resident frames only)"
    for rep in 1 2; do
        timeout 90 python bench.py --quick --steps 40 --warmup 24 2>/dev/null | tail -1 | \
            python -c "import json,sys; d=json.loads(sys.stdin.read()); print('compact keys ms/frame', round(d['ms_per_step'],4), 'launches', d['gpu_launches'])" | tee -a gpurun_out/ab_keys.log
        PLS_GS_FULLKEYS=1 timeout 90 python bench.py --quick --steps 40 --warmup 24 2>/dev/null | tail -1 | \
            python -c "import json,sys; d=json.loads(sys.stdin.read()); print('full keys    ms/frame', round(d['ms_per_step'],4), 'launches', d['gpu_launches'])" | tee -a gpurun_out/ab_keys.log
    done ;;
later)
    echo "== warp-per-pending-normal in iterations >= 2 (PLS_KD_NGROUP_LATER=32): parity under the toggle, then A/B"
    PLS_KD_NGROUP_LATER=32 timeout 60 python -m pytest tests/test_gpu_parity.py tests/test_next_rows_gpu.py -q -m gpu --timeout 50 -p no:cacheprovider \
        -k "icp_kd_small or icp_cfg2 or nan_rows or reinit or whole_shipped_chain" > gpurun_out/pytest_later32.log 2>&1
    tail -4 gpurun_out/pytest_later32.log
    for rep in 1 2; do
        for later in 0 32; do
            PLS_KD_NGROUP_LATER=$later timeout 40 python bench.py --quick --steps 40 --warmup 24 2>/dev/null | tail -1 | \
                python -c "import json,sys; d=json.loads(sys.stdin.read()); print('later $later ms/frame', round(d['ms_per_step'],4), 'iters', d['iters_mean'])" \
                | tee -a gpurun_out/ab_later.log
        done
    done ;;
mgpu)
    # needs `gpurun --gpus N`: sharded == single-GPU poses for both maps and both exchange modes, then the quick bench
    # at 2 .. N ranks with the peer-to-peer and the NCCL exchange
    NG=$(python -c "import torch; print(torch.cuda.device_count())")
    echo "== multi-GPU on $NG GPUs"
    timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu --timeout 300 -rf > gpurun_out/pytest_mgpu.log 2>&1
    tail -5 gpurun_out/pytest_mgpu.log
    for n in 2 4 8; do
        [ "$n" -le "$NG" ] || continue
        for comm in p2p nccl; do
            timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 \
                bench.py --gpus $n --quick --steps 40 --warmup 24 --comm $comm 2>/dev/null | tail -1 | tee -a gpurun_out/mgpu_quick.log
        done
    done ;;
bench)
    echo "== bench"
    timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
    tail -c 600 gpurun_out/bench_n1.json ;;
esac
done
ls gpurun_out | head -30
