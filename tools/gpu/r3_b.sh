#!/bin/bash
mkdir -p gpurun_out/r3b
O=gpurun_out/r3b
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/summary.txt
tail -4 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt; tail -1 $O/smoke.log
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --no-c5 ) > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" >> $O/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3b/bench_n1.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ['value','ms_per_step','l2_call_ms','hl_batch_ms']}, d['e2e']['value'], d['e2e']['step_ms'])
PY
cat $O/summary.txt
