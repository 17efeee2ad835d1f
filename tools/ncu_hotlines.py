#!/usr/bin/env python
"""Per-source-line warp-stall samples of an ncu report captured with --import-source on.

    python tools/ncu_hotlines.py report.ncu-rep [--kernel SUBSTR] [--top 40]

For a latency-bound kernel the interesting lines are those where the *busy* warp sits
(samples excluding stall_barrier: warps parked at __syncthreads are waiting for that warp).
"""
import argparse
import collections
import csv
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--kernel", default="")
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = csv.reader(out.splitlines())
    fpath = func = None
    hdr = None
    per = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0, 0, ""]))
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            func = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_all = hdr.index("Warp Stall Sampling (All Samples)")
            i_bar = hdr.index("stall_barrier")
            i_exec = hdr.index("Instructions Executed")
            continue
        if hdr is None or not r[0] or not r[0].isdigit():
            continue
        if a.kernel and a.kernel not in func:
            continue
        try:
            s_all, s_bar, ex = int(r[i_all]), int(r[i_bar]), int(r[i_exec])
        except ValueError:
            continue
        e = per[func][(fpath, int(r[0]))]
        e[0] += s_all
        e[1] += s_bar
        e[2] += ex
        e[3] = r[1].strip()[:110]
    for func, lines in per.items():
        tot = sum(v[0] for v in lines.values())
        bar = sum(v[1] for v in lines.values())
        print(f"== {func[:90]}\n   samples {tot}, at barriers {bar} ({100.0 * bar / max(tot, 1):.1f} %), busy {tot - bar}")
        top = sorted(lines.items(), key=lambda kv: -(kv[1][0] - kv[1][1]))[:a.top]
        for (f, ln), v in top:
            print(f"   {v[0] - v[1]:7d} busy {v[1]:7d} bar {v[2]:9d} inst  {f}:{ln:<5d} {v[3]}")


if __name__ == "__main__":
    main()
