#!/usr/bin/env python
"""Profile searches one call at a time against iadh_profile_search_batch on a bench workload (wall + device ms).

    python tools/hl_batch_probe.py [--workload c4] [--profiles 32] [--reps 3]
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
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--profiles", type=int, default=32)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-seq", action="store_true")
    a = ap.parse_args()
    cfg = bench.workload_config(a.workload)
    ds = ingest.dataset_from_synth(cfg)
    eng = capi.load_product(ds)
    eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    bp, src = bench.bench_profiles(ds, a.workload, a.profiles, lambda: eng.level2_all_pairs().bytes)
    print(src, file=sys.stderr)
    profs = [(None, p) for _, p in bp]
    masks = [m for m, _ in bp]
    for rep in range(a.reps):
        eng.clear_masks()
        t0 = time.perf_counter()
        seq = []
        for (m, p), mk in zip(profs, masks):
            if a.no_seq:
                break
            for (l, b, e) in mk:
                eng.mask_range(l, b, e)
            seq.append(eng.profile_search(p))
        t1 = time.perf_counter()
        eng.clear_masks()
        t2 = time.perf_counter()
        bat = eng.profile_search_batch([p for _, p in profs], masks)
        t3 = time.perf_counter()
        same = all(x.bytes == y.bytes for x, y in zip(seq, bat)) if seq else None
        print(f"rep {rep}: sequential {1e3 * (t1 - t0):.2f} ms wall ({sum(s.stats['total_ms'] for s in seq):.2f} device), "
              f"batch {1e3 * (t3 - t2):.2f} ms wall ({sum(s.stats['total_ms'] for s in bat):.2f} device, cluster "
              f"{sum(s.stats['cluster_ms'] for s in bat):.2f}, build {sum(s.stats['build_ms'] for s in bat):.2f}), "
              f"GHMs {sum(s.stats['n_ghm'] for s in bat)}, same bytes: {same}", file=sys.stderr)
    eng.close()


if __name__ == "__main__":
    main()
