#!/bin/bash
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/summary.txt
tail -4 $O/pytest.log
IADH_PASS_TIMING=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 4 --reps 3 > $O/trace_c4.log 2>&1
grep "pass timing" $O/trace_c4.log | tail -1
grep "call wall" $O/trace_c4.log | tail -5 | sed 's/stats.*build_ms/build_ms/' | cut -c1-150
IADH_RESIDENT=1 IADH_PASS_TIMING=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 4 --reps 2 > $O/trace_c4_res.log 2>&1
grep -A1 "resident launch" $O/trace_c4_res.log | head -16 | cut -c1-200
timeout 600 python tools/trace_run.py --workload c3 --profiles 4 --reps 3 > $O/trace_c3.log 2>&1
grep "call wall" $O/trace_c3.log | tail -5 | sed 's/stats.*build_ms/build_ms/' | cut -c1-150
