#!/bin/bash
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/summary.txt
tail -4 $O/pytest.log
timeout 600 python tools/hl_batch_probe.py --workload c4 --profiles 32 --reps 3 > $O/hlb_c4.log 2>&1
tail -3 $O/hlb_c4.log
timeout 600 python tools/hl_batch_probe.py --workload c3 --profiles 32 --reps 3 > $O/hlb_c3.log 2>&1
tail -3 $O/hlb_c3.log
IADH_TRACE=1 timeout 600 python tools/hl_batch_probe.py --workload c4 --profiles 32 --reps 1 > $O/hlb_c4_trace.log 2>&1
grep "iadh\]" $O/hlb_c4_trace.log | tail -16
