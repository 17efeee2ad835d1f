#!/bin/bash
# Round 2, call R (2 GPUs): per-call wall times on both ranks, then the bench at N=2
mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dist_timing.py c4 > $O/timing.log 2>&1
grep "rank 0" $O/timing.log | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench n2 rc=$?" >> $O/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --no-c5 > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench n1 rc=$?" >> $O/summary.txt
python - <<'PY'
import json
for f in ['bench_n1','bench_n2']:
    d=json.loads(open('gpurun_out/r2r/%s.json'%f).read().strip().splitlines()[-1])
    print(f, {k: d.get(k) for k in ['value','ms_per_step','wall_ms_per_step','l2_call_ms','hl_batch_ms']}, d['e2e']['value'], d['roofline']['launch_ms'], d['roofline_build']['launch_ms'])
PY
cat $O/summary.txt
