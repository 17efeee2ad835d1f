#!/usr/bin/env python
"""One profile search (optionally restricted to one y-list with IADH_ONLY_LIST) for critical-path studies.

    IADH_TRACE=1 python tools/one_hl.py [--profile 0] [--reps 3]
"""
import argparse, os, sys
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python")); sys.path.insert(0, REPO)
from iadhore_b200 import capi, ingest
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--profile", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
cfg = bench.workload_config("c3")
ds = ingest.dataset_from_synth(cfg)
only = os.environ.pop("IADH_ONLY_LIST", None)
eng = capi.load_product(ds)
eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
r = eng.level2_all_pairs()
profs = bench.make_profiles(ds, r.bytes, a.profile + 1)
m, p = profs[a.profile]
eng.mask_range(m.x_object_id, m.bounds[0], m.bounds[1]); eng.mask_range(m.y_object_id, m.bounds[2], m.bounds[3])
if only is not None:
    if only == "heaviest":
        import numpy as np
        dots = eng.build_dots([0] * ds.n_lists, list(range(ds.n_lists)), profile=p)
        only = str(int(np.argmax([len(a0) + len(a1) for a0, a1 in dots])))
        print("heaviest list:", only, file=sys.stderr)
    os.environ["IADH_ONLY_LIST"] = only
for _ in range(a.reps):
    pr = eng.profile_search(p)
    print(f"profile {a.profile} S={p.S} L={p.L}: {pr.stats['n_ghm']} GHMs {pr.stats['n_dots']} dots cluster {pr.stats['cluster_ms']:.3f} ms total {pr.stats['total_ms']:.3f} ms", file=sys.stderr)
eng.close()
