#!/bin/bash
# Round 2, call Y: mid-size profile batches (what a rank sees at 4 / 8 GPUs): K2 variants
mkdir -p gpurun_out/r2y
O=gpurun_out/r2y
for P in 8 4; do
  for V in "default:" "nores:IADH_RESIDENT=0" "res1:IADH_RESIDENT=1" "pass:IADH_PASS_MIN=500 IADH_PASS_RATIO=0 IADH_RESIDENT=0" "narrow:IADH_RESIDENT=0 IADH_CLASSES=1000:256:2,0:64:8" "narrow32:IADH_RESIDENT=0 IADH_CLASSES=1000:256:2,0:32:16" "three:IADH_RESIDENT=0 IADH_CLASSES=1200:256:1,500:128:2,0:32:16"; do
    name=${V%%:*}; envs=${V#*:}
    env $envs timeout 600 python tools/hl_batch_probe.py --workload c4 --profiles $P --reps 3 --no-seq > $O/p${P}_$name.log 2>&1
    echo "P=$P $name: $(tail -1 $O/p${P}_$name.log | cut -c40-140)"
  done
done
IADH_TRACE=1 timeout 600 python tools/hl_batch_probe.py --workload c4 --profiles 4 --reps 2 --no-seq > $O/p4_trace.log 2>&1
grep "iadh\]" $O/p4_trace.log | tail -24 | cut -c1-200
