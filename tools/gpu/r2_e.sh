#!/bin/bash
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
for kb in 4 24 32 48; do
  IADH_JOIN2=1 IADH_RES_MAX_KB=$kb timeout 600 python tools/trace_run.py --workload c4 --profiles 0 --reps 3 > $O/hyb_$kb.log 2>&1
  echo "== res_max_kb=$kb"; grep "L2 call wall" $O/hyb_$kb.log | tail -1 | sed 's/stats.*build_ms/build_ms/' | cut -c1-160
done
IADH_JOIN2=1 IADH_RES_MAX_KB=24 IADH_PASS_TIMING=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 0 --reps 2 > $O/hyb_24_timing.log 2>&1
grep "pass timing" $O/hyb_24_timing.log | tail -1
