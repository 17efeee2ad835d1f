#!/usr/bin/env python
"""Run the K2 pipeline on ONE list pair (critical-path studies: with a single GHM in flight the
ncu stall samples of the clustering kernel are exactly that GHM's serial latency).

    python tools/one_ghm.py [--workload c3] [--index 768 | --heaviest] [--reps 3]
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
    ap.add_argument("--index", type=int, default=-1, help="position in the level-2 pair sequence (-1: most dots)")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    cfg = bench.workload_config(a.workload)
    ds = ingest.dataset_from_synth(cfg)
    eng = capi.load_product(ds)
    eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    xs, ys = capi.plan_pairs(np.diff(ds.list_off))
    idx = a.index
    if idx < 0:
        dots = eng.build_dots(xs, ys)
        idx = int(np.argmax([len(d0) + len(d1) for d0, d1 in dots]))
    for _ in range(a.reps):
        r = eng.level2_pairs([xs[idx]], [ys[idx]])
        print(f"pair #{idx} ({xs[idx]},{ys[idx]}): {r.stats['n_dots']} dots, {r.n_multiplicons} multiplicons, "
              f"cluster {r.stats['cluster_ms']:.3f} ms, call {r.stats['total_ms']:.3f} ms", file=sys.stderr)
    eng.close()


if __name__ == "__main__":
    main()
