#!/usr/bin/env python
"""Per-launch DRAM bytes and duration of the kernels of ONE level-2 call, from
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv`
over `tools/trace_run.py --profiles 0 --reps R` (the last call's launches are kept).

    python tools/ncu_dram.py raw.csv out.csv [traffic.json]

out.csv: kernel, read bytes, written bytes, ns per launch; traffic.json (profiles/ncu_traffic.json, read by
bench.py for roofline.traffic): the sums per stage (K1 expansion path, K2 pass-per-launch path)."""
import collections
import csv
import json
import os
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    ii, ki, mi, vi, ui = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].replace("void <unnamed>::", "").replace("<unnamed>::", "").split("(")[0]
        d = launches.setdefault(int(r[ii]), {"kernel": name})
        d[r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
    seq = list(launches.values())
    # the last level-2 call: from the last init_tasks_kernel on
    start = max(i for i, d in enumerate(seq) if d["kernel"].startswith("init_tasks_kernel"))
    call = seq[start:]
    with open(dst, "w") as fh:
        fh.write("# ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none: "
                 "every launch of one level-2 call on config 4 (bench workload), in launch order\n")
        fh.write("kernel,dram_read_bytes,dram_write_bytes,duration_ns\n")
        for d in call:
            fh.write(f"\"{d['kernel']}\",{int(d.get('dram__bytes_read.sum', 0))},{int(d.get('dram__bytes_write.sum', 0))},"
                     f"{int(d.get('gpu__time_duration.sum', 0))}\n")
    if len(sys.argv) > 3:
        tot = lambda pred: int(sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in call if pred(d["kernel"])))
        cur = json.load(open(sys.argv[3])) if os.path.exists(sys.argv[3]) else {}
        cur["K2 pass-per-launch, config 4 level-2 call (phase1_pass x20 + phase1_join2 x10 + phase2_stage x4 launches)"] = \
            tot(lambda k: k.startswith("phase1_") or k.startswith("phase2_"))
        cur["K1 expansion, config 4 level-2 call (l2_expand count + fill, block sorts)"] = \
            tot(lambda k: k.startswith("l2_expand") or k.startswith("seg_sort") or k.startswith("scatter_cnt"))
        json.dump(cur, open(sys.argv[3], "w"), indent=1, sort_keys=True)
    t = collections.Counter()
    for d in call:
        t[d["kernel"]] += d.get("gpu__time_duration.sum", 0)
    for k, v in t.most_common():
        print(f"{k:50s} {v / 1e6:8.3f} ms")


if __name__ == "__main__":
    main()
