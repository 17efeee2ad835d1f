#!/bin/bash
# shm exchange at 2 GPUs: correctness both ways, timing, bench
mkdir -p gpurun_out/r3a
O=gpurun_out/r3a
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py c3s0.3 > $O/dist_check_shm.log 2>&1
echo "dist_check shm rc=$?" >> $O/summary.txt
grep "rank" $O/dist_check_shm.log | tail -8 | cut -c1-200
IADH_SHM_EXCHANGE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py c3s0.3 > $O/dist_check_nccl.log 2>&1
echo "dist_check nccl rc=$?" >> $O/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dist_timing.py c4 > $O/timing.log 2>&1
grep "rep [34]" $O/timing.log | tr ']' '\n' | grep "rep" | sort
bash tools/gpu/scale.sh 2
ls /dev/shm | head
cat $O/summary.txt
