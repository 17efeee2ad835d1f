#!/bin/bash
# Round-end evidence, run on the GPU box (gpurun): launch list of one bench run, full captures of
# the dominant kernels.  Usage: bash tools/profile_round.sh <tag>   (outputs under gpurun_out/)
set -x
TAG=${1:-r01}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c5 > gpurun_out/${TAG}_bench_for_ncu.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_raw.csv \
    python bench.py --steps 1 --warmup 3 --profiles 8 --no-cpu-baseline --no-c5 > gpurun_out/${TAG}_ncu_launches.log 2>&1
# K2 of the level-2 call (the three size-class launches of the second L2 call) and of one profile search,
# on the bench workload (config 4)
ncu --set full --clock-control none --import-source on -k regex:cluster_kernel --launch-skip 3 --launch-count 6 \
    -o gpurun_out/${TAG}_cluster -f python tools/trace_run.py --workload c4 --profiles 1 --reps 2 > gpurun_out/${TAG}_ncu_cluster.log 2>&1
# K1 on config 5 (dense) : count + fill
ncu --set full --clock-control none --import-source on -k regex:l2_build_kernel --launch-skip 2 --launch-count 2 \
    -o gpurun_out/${TAG}_build_c5 -f python tools/build_probe.py --workload c5 --reps 2 > gpurun_out/${TAG}_ncu_build.log 2>&1
# K3
ncu --set full --clock-control none -k regex:pack_ --launch-skip 2 --launch-count 2 \
    -o gpurun_out/${TAG}_pack -f python tools/trace_run.py --workload c4 --profiles 0 --reps 2 > gpurun_out/${TAG}_ncu_pack.log 2>&1
ls -la gpurun_out/
