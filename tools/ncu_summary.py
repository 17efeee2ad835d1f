"""Summarise ncu output brought back in gpurun_out/ into small text files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.csv
    python tools/ncu_summary.py report   gpurun_out/prof_cluster.ncu-rep profiles/r01_cluster_kernel.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    ki, mi, vi, ui = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as fh:
        fh.write("# per-kernel aggregate of `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised)\n")
        fh.write("kernel,launches,total_us,share,max_us\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"{k},{a[0]},{a[1]:.1f},{a[1] / tot:.4f},{a[2]:.1f}\n")


def report(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as fh:
        fh.write(f"# selected metrics of {src} (ncu --set full --clock-control none), one block per captured launch\n")
        for r in rows[2:]:
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    fh.write(f"{k:90s} {r[i]} {units[i]}\n")
            fh.write("\n")


def traffic(src, dst):
    """Merge dram__bytes_read.sum + dram__bytes_write.sum per launch (max over the captured launches
    of each kernel) of an `ncu --set full` report into the JSON bench.py reads (roofline.traffic)."""
    import json
    import os
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    cur = json.load(open(dst)) if os.path.exists(dst) else {}
    for r in rows[2:]:
        name = r[ik].replace("void <unnamed>::", "").split("(")[0]
        b = float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
        cur[name] = max(cur.get(name, 0), int(b)) if os.environ.get("NCU_TRAFFIC_KEEP_MAX") else int(b)
    json.dump(cur, open(dst, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    {"launches": launches, "report": report, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
