#!/bin/bash
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py c3s0.3 > $O/dist_check.log 2>&1
echo "dist_check rc=$?" >> $O/summary.txt
grep "rank" $O/dist_check.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench n2 rc=$?" >> $O/summary.txt
tail -c 1500 $O/bench_n2.json; tail -3 $O/bench_n2.err
cat $O/summary.txt
