#!/bin/bash
O=gpurun_out/r3c; mkdir -p $O
for V in "default:" "resident:IADH_RESIDENT=1" "fused:IADH_PASS_MIN=100000000 IADH_RESIDENT=0" "res_fused:IADH_PASS_MIN=100000000 IADH_RESIDENT=1" "cfg3224:IADH_PASS_CFG=32:24"; do
  name=${V%%:*}; envs=${V#*:}
  echo "== $name"
  env $envs timeout 300 python tools/l2_share_probe.py c4 8 4 2>&1 | grep share
done
