export IADH_CLASSES=0:128:4
IADH_TRACE=1 python tools/one_ghm.py --index 768 --reps 2 2>&1 | grep -E "slowest|join of" | tail -2
python tools/sweep_classes.py --sets "1025:256:2,193:128:4,0:32:16;0:128:4;0:128:5" | tail -3
