#!/bin/bash
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/summary.txt
tail -4 $O/pytest.log
( time timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --no-c5 ) > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" >> $O/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n/bench_n1.json').read().strip().splitlines()[-1])
for k in ['value','ms_per_step','wall_ms_per_step','l2_call_ms','hl_batch_ms','hl_single_call_ms']: print(k, d.get(k))
print(d['e2e'])
PY
cat $O/summary.txt
