#!/bin/bash
# Round 2, call V: workspace layout in queue order vs caller's order: per-pass timing on config 4, HL batch, parity subset
mkdir -p gpurun_out/r2v
O=gpurun_out/r2v
IADH_PASS_TIMING=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 0 --reps 3 > $O/queue_order.log 2>&1
grep "pass timing" $O/queue_order.log | tail -1
grep "L2 call wall" $O/queue_order.log | tail -1 | cut -c1-200
IADH_WS_ORDER=0 IADH_PASS_TIMING=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 0 --reps 3 > $O/caller_order.log 2>&1
grep "pass timing" $O/caller_order.log | tail -1
grep "L2 call wall" $O/caller_order.log | tail -1 | cut -c1-200
timeout 600 python tools/hl_batch_probe.py --workload c4 --profiles 32 --reps 3 --no-seq > $O/hl_queue.log 2>&1; tail -1 $O/hl_queue.log | cut -c1-200
IADH_WS_ORDER=0 timeout 600 python tools/hl_batch_probe.py --workload c4 --profiles 32 --reps 3 --no-seq > $O/hl_caller.log 2>&1; tail -1 $O/hl_caller.log | cut -c1-200
( timeout 900 python -m pytest tests/test_gpu_bench_parity.py tests/test_gpu_parity.py -m gpu -x -q -k "bench or c3 or pass_per_launch or small" ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/summary.txt; tail -2 $O/pytest.log
