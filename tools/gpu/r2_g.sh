#!/bin/bash
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/summary.txt
tail -4 $O/pytest.log
IADH_PASS_TIMING=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 4 --reps 3 > $O/trace_c4.log 2>&1
grep "pass timing" $O/trace_c4.log | tail -1
grep "call wall" $O/trace_c4.log | tail -5 | sed 's/stats.*build_ms/build_ms/' | cut -c1-150
