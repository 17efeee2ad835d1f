#!/bin/bash
N=${1:-4}
O=gpurun_out/scale
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/dist_timing.py c4 > $O/timing_n$N.log 2>&1
grep "rep [67]" $O/timing_n$N.log | tr '[' '\n' | grep "rep" | sort | head -16
bash tools/gpu/scale.sh $N
