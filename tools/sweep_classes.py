#!/usr/bin/env python
"""Sweep of the K2 size-class thresholds (IADH_CLASSES, read by iadh_create) and
other tuning knobs on one workload: prints the clustering-kernel time of the level-2 call and
the mean profile-search call time.

    python tools/sweep_classes.py [--workload c3] [--profiles 8] [--sets "192:1024,64:1024,..."]
"""
import argparse
import os
import sys

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python"))
sys.path.insert(0, REPO)
from iadhore_b200 import capi, ingest  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--profiles", type=int, default=8)
    ap.add_argument("--sets", default="1025:256:2,193:128:4,0:32:16;0:128:4;0:128:5;0:64:8;0:64:10;1500:128:2,0:64:8;1500:256:1,0:64:8",
                    help="semicolon-separated IADH_CLASSES values")
    ap.add_argument("--env", default="", help="extra VAR=VALUE;VAR=VALUE applied to every run")
    a = ap.parse_args()
    cfg = bench.workload_config(a.workload)
    ds = ingest.dataset_from_synth(cfg)
    for kv in a.env.split(";"):
        if kv:
            k, v = kv.split("=")
            os.environ[k] = v
    profs = None
    for s in a.sets.split(";"):
        os.environ["IADH_CLASSES"] = s
        eng = capi.load_product(ds)
        eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
        l2 = []
        for _ in range(3):
            r = eng.level2_all_pairs()
            l2.append((r.stats["cluster_ms"], r.stats["total_ms"]))
        if profs is None:
            profs = bench.make_profiles(ds, r.bytes, a.profiles)
        hl = []
        for rep in range(2):
            eng.clear_masks()
            for m, p in profs:
                eng.mask_range(m.x_object_id, m.bounds[0], m.bounds[1])
                eng.mask_range(m.y_object_id, m.bounds[2], m.bounds[3])
                pr = eng.profile_search(p)
                if rep:
                    hl.append((pr.stats["cluster_ms"], pr.stats["total_ms"]))
        eng.close()
        l2 = np.array(l2[1:])
        hl = np.array(hl)
        print(f"{s:36s} L2 cluster {l2[:, 0].mean():7.2f} ms total {l2[:, 1].mean():7.2f} ms | "
              f"HL cluster {hl[:, 0].mean():6.2f} ms total {hl[:, 1].mean():6.2f} ms (max {hl[:, 1].max():.2f})", flush=True)


if __name__ == "__main__":
    main()
