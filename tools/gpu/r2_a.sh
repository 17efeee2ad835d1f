#!/bin/bash
# Round 2, GPU call A: sanity of the suite, JOIN2 on a device for the first time, per-pass timing on config 4.
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_default.log 2>&1
echo "pytest default rc=$?" >> $O/summary.txt
( time IADH_JOIN2=1 IADH_PASS_MIN=1 IADH_PASS_RATIO=0 timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_join2.log 2>&1
echo "pytest join2 forced-pass rc=$?" >> $O/summary.txt
IADH_PASS_TIMING=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 2 --reps 3 > $O/trace_c4.log 2>&1
echo "trace c4 rc=$?" >> $O/summary.txt
IADH_JOIN2=1 IADH_PASS_TIMING=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 2 --reps 3 > $O/trace_c4_join2.log 2>&1
echo "trace c4 join2 rc=$?" >> $O/summary.txt
IADH_TRACE=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 2 --reps 2 > $O/trace_c4_stages.log 2>&1
echo "trace c4 stages rc=$?" >> $O/summary.txt
cat $O/summary.txt
tail -3 $O/pytest_default.log $O/pytest_join2.log
grep "pass timing" $O/trace_c4.log | tail -1
grep "pass timing" $O/trace_c4_join2.log | tail -1
