#!/bin/bash
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/summary.txt
tail -4 $O/pytest.log
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" >> $O/summary.txt
tail -c 3000 $O/bench_n1.json; tail -5 $O/bench_n1.err
