#!/bin/bash
# Round 2, call Q: K1 expansion for profile GHMs + register sort: parity, bench, traces, K1 ncu
mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/summary.txt
tail -4 $O/pytest.log
( time timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra ) > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" >> $O/summary.txt
IADH_TRACE=1 timeout 600 python tools/hl_batch_probe.py --workload c4 --profiles 32 --reps 2 > $O/hl_trace.log 2>&1
grep "iadh\]\|^rep" $O/hl_trace.log | tail -60 | cut -c1-150
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"l2_expand|seg_sort" --launch-skip 4 --launch-count 4 \
    -o $O/k1_expand -f python tools/trace_run.py --workload c4 --profiles 0 --reps 2 > $O/ncu_k1.log 2>&1
echo "ncu k1 rc=$?" >> $O/summary.txt
python tools/ncu_summary.py report $O/k1_expand.ncu-rep $O/k1_expand.txt
rm -f $O/k1_expand.ncu-rep
cat $O/summary.txt
