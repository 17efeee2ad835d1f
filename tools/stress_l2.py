#!/usr/bin/env python
"""Repeat the level-2 call of a workload and compare every result with the first (race hunting)."""
import argparse, hashlib, os, sys
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python")); sys.path.insert(0, REPO)
from iadhore_b200 import capi, ingest
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4")
ap.add_argument("--reps", type=int, default=6)
a = ap.parse_args()
cfg = bench.workload_config(a.workload)
ds = ingest.dataset_from_synth(cfg)
eng = capi.load_product(ds)
eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
first = None
for i in range(a.reps):
    r = eng.level2_all_pairs()
    h = hashlib.sha1(r.bytes).hexdigest()
    first = first or h
    print(i, r.stats["n_ghm"], r.stats["n_dots"], r.n_multiplicons, h[:12], "OK" if h == first else "DIFFERENT", flush=True)
