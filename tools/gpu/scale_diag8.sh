#!/bin/bash
O=gpurun_out/scale
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 tools/dist_timing.py c4 > $O/timing_n8.log 2>&1
grep "rep 4" $O/timing_n8.log | tr ']' '\n' | grep "rep 4" | sort | head -10
nproc; free -g | head -2
