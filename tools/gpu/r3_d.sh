#!/bin/bash
for V in "3232_32:IADH_PASS_CFG=3232:32" "3224_24:IADH_PASS_CFG=3224:24" "32_16:IADH_PASS_CFG=32:16" "32_24:IADH_PASS_CFG=32:24" "32_32:IADH_PASS_CFG=32:32" "6416_16:IADH_PASS_CFG=6416:16" "64_8:IADH_PASS_CFG=64:8"; do
  name=${V%%:*}; envs=${V#*:}
  echo "== $name"
  env $envs timeout 300 python tools/l2_share_probe.py c4 16 8 4 1 2>&1 | grep share
done
