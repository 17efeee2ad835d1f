for c in ${CLASSES:-0:64:8 0:128:4 0:128:5 0:256:2 0:128:8}; do
  export IADH_CLASSES=$c
  echo "== $c"
  IADH_TRACE=1 python tools/trace_run.py --profiles 0 --reps 2 2>&1 | grep -E "K2 cluster|all GHMs|slowest" | tail -3
done
