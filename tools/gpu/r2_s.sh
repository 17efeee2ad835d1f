#!/bin/bash
# Round 2, call S (2 GPUs): zero-copy batch results / gathers: parity suite, exchange check, per-call wall times, bench at N=1,2
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/summary.txt
tail -3 $O/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py c3s0.3 > $O/dist_check.log 2>&1
echo "dist_check rc=$?" >> $O/summary.txt
grep "rank" $O/dist_check.log | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dist_timing.py c4 > $O/timing.log 2>&1
grep "rep [34]" $O/timing.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench n2 rc=$?" >> $O/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --no-c5 > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench n1 rc=$?" >> $O/summary.txt
python - <<'PY'
import json
for f in ['bench_n1','bench_n2']:
    d=json.loads(open('gpurun_out/r2s/%s.json'%f).read().strip().splitlines()[-1])
    print(f, {k: d.get(k) for k in ['value','ms_per_step','wall_ms_per_step','l2_call_ms','hl_batch_ms']}, d['e2e']['value'], d['roofline']['launch_ms'], d['roofline_build']['launch_ms'])
PY
cat $O/summary.txt
