#!/bin/bash
# Round 2, call P: K1 family-expansion path: parity, bench, stage trace; fresh ncu evidence summarised on the box
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/summary.txt
tail -4 $O/pytest.log
( time timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra ) > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" >> $O/summary.txt
IADH_TRACE=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 0 --reps 3 > $O/trace_c4.log 2>&1
grep "iadh\]" $O/trace_c4.log | tail -32 | cut -c1-150
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_raw.csv \
    python bench.py --steps 1 --warmup 3 --profiles 8 --no-cpu-baseline --no-c5 --no-extra > $O/ncu_launches.log 2>&1
echo "ncu launches rc=$?" >> $O/summary.txt
python tools/ncu_summary.py launches $O/launches_raw.csv $O/launches.csv
# second L2 call of trace_run: phase-1 launches 30..59; gap index 4 = launches 42,43,44
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase1_ --launch-skip 42 --launch-count 3 \
    -o $O/pass_g4 -f python tools/trace_run.py --workload c4 --profiles 0 --reps 2 > $O/ncu_pass.log 2>&1
echo "ncu pass rc=$?" >> $O/summary.txt
python tools/ncu_summary.py report $O/pass_g4.ncu-rep $O/pass_g4.txt
python tools/ncu_hotlines.py $O/pass_g4.ncu-rep --top 50 > $O/pass_g4_hot.txt 2>&1
rm -f $O/pass_g4.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"l2_expand|seg_sort" --launch-skip 3 --launch-count 3 \
    -o $O/k1_expand -f python tools/trace_run.py --workload c4 --profiles 0 --reps 2 > $O/ncu_k1.log 2>&1
echo "ncu k1 rc=$?" >> $O/summary.txt
python tools/ncu_summary.py report $O/k1_expand.ncu-rep $O/k1_expand.txt
python tools/ncu_hotlines.py $O/k1_expand.ncu-rep --top 30 > $O/k1_expand_hot.txt 2>&1
rm -f $O/k1_expand.ncu-rep
cat $O/summary.txt
du -sh $O
