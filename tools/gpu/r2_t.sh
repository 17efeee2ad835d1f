#!/bin/bash
# Round 2, call T: stage trace of the level-2 call after the host-side changes; ncu of the first seed pass (7 ms); interning test
mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "interning or expansion" ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/summary.txt
tail -2 $O/pytest.log
IADH_TRACE=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 0 --reps 3 > $O/trace_c4.log 2>&1
grep "iadh\]" $O/trace_c4.log | tail -30 | cut -c1-150
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase1_pass --launch-skip 20 --launch-count 1 \
    -o $O/pass_seed0 -f python tools/trace_run.py --workload c4 --profiles 0 --reps 2 > $O/ncu_seed0.log 2>&1
echo "ncu seed0 rc=$?" >> $O/summary.txt
python tools/ncu_summary.py report $O/pass_seed0.ncu-rep $O/pass_seed0.txt
python tools/ncu_hotlines.py $O/pass_seed0.ncu-rep --top 60 > $O/pass_seed0_hot.txt 2>&1
rm -f $O/pass_seed0.ncu-rep
cat $O/summary.txt
