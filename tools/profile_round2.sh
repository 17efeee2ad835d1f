#!/bin/bash
# Round-2 evidence, run on the GPU box (gpurun): bench line, launch list of the same command, per-launch DRAM bytes of
# one level-2 call, full captures of the dominant kernels summarised on the box (the .ncu-rep files stay there).
# Usage: bash tools/profile_round2.sh   (outputs under gpurun_out/r02/)
O=gpurun_out/r02
mkdir -p $O
python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err || exit 1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c5 --no-extra > $O/bench_for_ncu.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c5 --no-extra > $O/ncu_launches.log 2>&1
python tools/ncu_summary.py launches $O/launches_raw.csv $O/launches.csv
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/dram_raw.csv \
    python tools/trace_run.py --workload c4 --profiles 0 --reps 2 > $O/ncu_dram.log 2>&1
python tools/ncu_dram.py $O/dram_raw.csv $O/l2_call_dram.csv $O/ncu_traffic.json > $O/l2_call_kernels.txt
cap() {  # name, kernel regex, skip, count, command...
    local name=$1 re=$2 skip=$3 cnt=$4; shift 4
    ncu --set full --clock-control none --import-source on -k regex:"$re" --launch-skip $skip --launch-count $cnt -o $O/$name -f "$@" > $O/ncu_$name.log 2>&1
    python tools/ncu_summary.py report $O/$name.ncu-rep $O/$name.txt
    python tools/ncu_hotlines.py $O/$name.ncu-rep --top 40 > $O/${name}_hot.txt 2>&1
    NCU_TRAFFIC_KEEP_MAX=1 python tools/ncu_summary.py traffic $O/$name.ncu-rep $O/ncu_traffic.json
    rm -f $O/$name.ncu-rep
}
# second level-2 call of trace_run: phase1_pass launches 20..39 (seed, enrich per gap), join2 launches 10..19
cap pass_seed0 "phase1_pass" 20 1 python tools/trace_run.py --workload c4 --profiles 0 --reps 2
cap pass_g4 "phase1_" 42 3 python tools/trace_run.py --workload c4 --profiles 0 --reps 2
cap k1_expand "l2_expand|seg_sort" 5 5 python tools/trace_run.py --workload c4 --profiles 0 --reps 2
cap k1_c5 "l2_build_kernel" 2 2 python tools/build_probe.py --workload c5 --reps 2
cap hl_cluster "cluster_kernel|hl_expand" 0 4 python tools/hl_batch_probe.py --workload c4 --profiles 8 --reps 1 --no-seq
# K4 (homolog-count filter): the probe first (plain vs filtered batch, sha1 of the streams), then the link-table kernels
# and the filter kernel of the filtered batch (not captured in round 2: the GPU budget ended with the probe)
python tools/hl_filter_probe.py --reps 3 --out $O/hl_filter_probe.json 2> $O/hl_filter_probe.err
cap k4_filter "link_|homolog_filter" 0 5 python tools/hl_filter_probe.py --reps 0
cuobjdump -sass i-adhore_b200/csrc/libiadhore_b200.so 2>/dev/null | grep -E "Function :|LDG\.E\.128|STG\.E\.128|ATOMG|RED\.|SHFL|LDGSTS|UTMA" | \
    awk '/Function :/ {f=$3} !/Function :/ {split($0,a," "); for(i in a) if (a[i] ~ /^(LDG|STG|ATOMG|RED|SHFL|LDGSTS|UTMA)/) c[f" "a[i]]++} END {for (k in c) print c[k], k}' | sort -k2,2 -k1,1nr > $O/sass_mnemonics.txt
ls -la $O
