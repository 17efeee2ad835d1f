#!/usr/bin/env python
"""Device time of one rank's share of the level-2 call (rank 0 of N) on one GPU -- what a rank of an N-GPU run computes --
for tuning the K2 path choice at small shares.   python tools/l2_share_probe.py [c4] [N ...]"""
import os, sys, time
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python")); sys.path.insert(0, REPO)
from iadhore_b200 import capi, ingest
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
ns = [int(x) for x in sys.argv[2:]] or [8, 4, 2]
cfg = bench.workload_config(wl)
ds = ingest.dataset_from_synth(cfg, device=0)
eng = capi.load_product(ds)
eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
for n in ns:
    best = None
    for rep in range(4):
        r = eng.level2_all_pairs(0, n)
        s = r.stats
        if best is None or s["total_ms"] < best["total_ms"]:
            best = s
    print(f"share 1/{n}: {best['n_ghm']} GHMs, call {best['total_ms']:.2f} ms, K2 {best['cluster_ms']:.2f} ms, K1 {best['build_ms']:.2f} ms, launches {best['kernel_launches']}", flush=True)
eng.close()
