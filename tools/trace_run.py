#!/usr/bin/env python
"""Stage-by-stage timing of the hot path (IADH_TRACE=1 makes the engine synchronise after every
stage and print its wall time + the per-phase SM cycle counters of the clustering kernel).

    IADH_TRACE=1 python tools/trace_run.py [--workload c3] [--profiles 3] [--reps 2]
"""
import argparse
import os
import sys
import time

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python"))
sys.path.insert(0, REPO)
from iadhore_b200 import capi, ingest  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--profiles", type=int, default=3)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    cfg = bench.workload_config(a.workload)
    ds = ingest.dataset_from_synth(cfg)
    t0 = time.time()
    eng = capi.load_product(ds)
    print(f"create: {1e3 * (time.time() - t0):.1f} ms", file=sys.stderr)
    eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    for rep in range(a.reps):
        print(f"==== L2 rep {rep}", file=sys.stderr)
        t0 = time.time()
        r = eng.level2_all_pairs()
        print(f"L2 call wall {1e3 * (time.time() - t0):.2f} ms; stats {r.stats}", file=sys.stderr)
    profs = bench.make_profiles(ds, r.bytes, a.profiles)
    for rep in range(a.reps):
        eng.clear_masks()
        for k, (m, p) in enumerate(profs):
            eng.mask_range(m.x_object_id, m.bounds[0], m.bounds[1])
            eng.mask_range(m.y_object_id, m.bounds[2], m.bounds[3])
            print(f"==== HL rep {rep} profile {k} (S={p.S} L={p.L})", file=sys.stderr)
            t0 = time.time()
            pr = eng.profile_search(p)
            print(f"HL call wall {1e3 * (time.time() - t0):.2f} ms; stats {pr.stats}", file=sys.stderr)
    t0 = time.time()
    eng.close()
    print(f"destroy: {1e3 * (time.time() - t0):.1f} ms", file=sys.stderr)
    t0 = time.time()
    eng = capi.load_product(ds)
    print(f"create (2nd): {1e3 * (time.time() - t0):.1f} ms", file=sys.stderr)
    eng.close()


if __name__ == "__main__":
    main()
