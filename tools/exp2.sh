export IADH_CLASSES=0:128:4
for c in 48,48,48,48 16,16,16,16 8,8,8,8 16,48,48,48 48,16,48,48 48,48,16,48; do
  export IADH_COOP=$c
  echo "== coop $c"
  IADH_TRACE=1 python tools/one_ghm.py --index 768 --reps 2 2>&1 | grep -E "slowest|join of" | tail -2
done
