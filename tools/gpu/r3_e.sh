#!/bin/bash
mkdir -p gpurun_out/r3e
O=gpurun_out/r3e
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/summary.txt
tail -4 $O/pytest.log
timeout 300 python tools/l2_share_probe.py c4 8 4 2 1 2>&1 | grep share
cat $O/summary.txt
