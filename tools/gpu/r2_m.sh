#!/bin/bash
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dist_timing.py c4 > $O/timing.log 2>&1
grep "rank 0" $O/timing.log | tail -5
