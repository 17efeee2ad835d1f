#!/bin/bash
# Round 2, call U: device-built join tables in iadh_create: parity suite, create time, bench (e2e)
mkdir -p gpurun_out/r2u
O=gpurun_out/r2u
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/summary.txt
tail -3 $O/pytest.log
timeout 600 python tools/trace_run.py --workload c4 --profiles 0 --reps 1 2>&1 | grep -i "create" > $O/create.log
IADH_HOST_TABLES=1 timeout 600 python tools/trace_run.py --workload c4 --profiles 0 --reps 1 2>&1 | grep -i "create" >> $O/create.log
cat $O/create.log
( time timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra ) > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" >> $O/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2u/bench_n1.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ['value','ms_per_step','l2_call_ms','hl_batch_ms','hl_single_call_ms']}, d['e2e'], d['roofline_build_c5'])
PY
cat $O/summary.txt
