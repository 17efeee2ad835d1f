#!/bin/bash
# Round 2, GPU call B: resident K2 for the first time on a device: parity suite, then timing on config 4 / 3
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/summary.txt
tail -5 $O/pytest.log
IADH_TRACE=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 2 --reps 2 > $O/trace_c4_stages.log 2>&1
echo "trace c4 stages rc=$?" >> $O/summary.txt
timeout 600 python tools/trace_run.py --workload c4 --profiles 4 --reps 3 > $O/trace_c4.log 2>&1
echo "trace c4 rc=$?" >> $O/summary.txt
IADH_RESIDENT=0 IADH_JOIN2=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 4 --reps 3 > $O/trace_c4_nores.log 2>&1
timeout 600 python tools/trace_run.py --workload c3 --profiles 4 --reps 3 > $O/trace_c3.log 2>&1
echo "trace c3 rc=$?" >> $O/summary.txt
for w in "600:128" "300:128" "300:256" "150:64" "100000:32"; do
  IADH_RES_WIDE=$w timeout 600 python tools/trace_run.py --workload c4 --profiles 6 --reps 2 > $O/trace_c4_wide_${w/:/_}.log 2>&1
done
cat $O/summary.txt
grep -h "call wall" $O/trace_c4.log | tail -9
