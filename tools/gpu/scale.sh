#!/bin/bash
# bench.py at N GPUs of one box (torchrun), as the driver launches it.  Usage: bash tools/gpu/scale.sh N
N=${1:-2}
O=gpurun_out/scale
mkdir -p $O
nvidia-smi -L > $O/gpus_n$N.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "bench n$N rc=$?"
python - $N <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads(open('gpurun_out/scale/bench_n%s.json'%n).read().strip().splitlines()[-1])
print('N',n,{k: d.get(k) for k in ['value','ms_per_step','wall_ms_per_step','l2_call_ms','hl_batch_ms']}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'K2', d['roofline']['launch_ms'], 'K1', d['roofline_build']['launch_ms'])
PY
tail -3 $O/bench_n$N.err
