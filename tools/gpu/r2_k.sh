#!/bin/bash
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
timeout 600 python tools/hl_batch_probe.py --workload c4 --profiles 32 --reps 3 > $O/default.log 2>&1; tail -2 $O/default.log
IADH_PASS_RATIO=0 timeout 600 python tools/hl_batch_probe.py --workload c4 --profiles 32 --reps 3 --no-seq > $O/pass.log 2>&1; tail -1 $O/pass.log
IADH_PASS_MIN=100000000 timeout 600 python tools/hl_batch_probe.py --workload c4 --profiles 32 --reps 3 --no-seq > $O/fused.log 2>&1; tail -1 $O/fused.log
IADH_PASS_MIN=100000000 IADH_CLASSES="1200:256:2,300:128:4,0:64:8" timeout 600 python tools/hl_batch_probe.py --workload c4 --profiles 32 --reps 3 --no-seq > $O/fused3.log 2>&1; tail -1 $O/fused3.log
IADH_TRACE=1 timeout 600 python tools/hl_batch_probe.py --workload c4 --profiles 32 --reps 2 --no-seq > $O/trace.log 2>&1; grep "iadh\]" $O/trace.log | tail -22 | cut -c1-200
