for sm in 3 128 256 512 100000; do
  export IADH_SPLIT_MIN=$sm
  echo "== split_min $sm"
  python tools/sweep_classes.py --sets "1025:256:2,193:128:4,0:32:16" | tail -1
done
