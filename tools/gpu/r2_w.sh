#!/bin/bash
# Round 2, call W: profile searches: fused vs pass-per-launch for the batch, resident kernel for single calls
mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
timeout 600 python tools/hl_batch_probe.py --workload c4 --profiles 32 --reps 3 > $O/default.log 2>&1; tail -1 $O/default.log | cut -c1-220
IADH_PASS_RATIO_HL=100000000 timeout 600 python tools/hl_batch_probe.py --workload c4 --profiles 32 --reps 3 > $O/hl_fused.log 2>&1; tail -1 $O/hl_fused.log | cut -c1-220
IADH_RESIDENT=1 timeout 600 python tools/hl_batch_probe.py --workload c4 --profiles 32 --reps 3 > $O/resident.log 2>&1; tail -1 $O/resident.log | cut -c1-220
IADH_RESIDENT=1 IADH_RES_WIDE=300:128 timeout 600 python tools/hl_batch_probe.py --workload c4 --profiles 32 --reps 3 > $O/resident_wide.log 2>&1; tail -1 $O/resident_wide.log | cut -c1-220
IADH_PASS_RATIO_HL=100000000 timeout 600 python tools/hl_batch_probe.py --workload c3 --profiles 32 --reps 3 > $O/c3_fused.log 2>&1; tail -1 $O/c3_fused.log | cut -c1-220
timeout 600 python tools/hl_batch_probe.py --workload c3 --profiles 32 --reps 3 > $O/c3_default.log 2>&1; tail -1 $O/c3_default.log | cut -c1-220
