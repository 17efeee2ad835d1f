#!/usr/bin/env python
"""The profile-search batch of a bench workload with and without the device homolog-count filter
(iadh_profile_search_batch vs iadh_profile_search_batch_filtered): wall / device ms, multiplicons, result bytes.

    python tools/hl_filter_probe.py [--workload c4] [--profiles 32] [--reps 3] [--out gpurun_out/hf_probe.json]
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python"))
sys.path.insert(0, REPO)
from iadhore_b200 import capi, ingest, profiles  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--profiles", type=int, default=32)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    t0 = time.perf_counter()
    cfg = bench.workload_config(a.workload)
    ds = ingest.dataset_from_synth(cfg)
    path = bench.ref_profiles_path(a.workload)
    bp = profiles.load_reference_profiles(ds, path, a.profiles, links_path=path.replace("bench_profiles_", "bench_profile_links_").replace(".bin.gz", ".bin.xz"))
    profs = [p for _, p in bp]
    masks = [m for m, _ in bp]
    n_links = sum(len(p.links) for p in profs)
    eng = capi.load_product(ds)
    eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    setup_s = time.perf_counter() - t0
    out = {"workload": a.workload, "profiles": len(profs), "links": n_links, "setup_s": round(setup_s, 2), "plain": [], "filtered": []}
    for rep in range(a.reps + 1):   # (rep 0 = warm-up, not reported)
        for kind in ("plain", "filtered"):
            eng.clear_masks()
            t1 = time.perf_counter()
            res = eng.profile_search_batch(profs, masks) if kind == "plain" else eng.profile_search_batch_filtered(profs, masks)
            t2 = time.perf_counter()
            for r in res:
                r.wait()
            t3 = time.perf_counter()
            if rep:
                out[kind].append({"call_ms": round(1e3 * (t2 - t1), 3), "with_copy_ms": round(1e3 * (t3 - t1), 3),
                                  "device_ms": round(sum(r.stats["total_ms"] for r in res), 3),
                                  "cluster_ms": round(sum(r.stats["cluster_ms"] for r in res), 3),
                                  "multiplicons": sum(r.n_multiplicons for r in res), "bytes": sum(r._n for r in res),
                                  "ghms": sum(r.stats["n_ghm"] for r in res), "launches": sum(r.stats["kernel_launches"] for r in res),
                                  "h2d": sum(r.stats["h2d_bytes"] for r in res)})
    # streams of the last repetition, for comparison with the oracle's (profiles/parity_r02.md)
    import hashlib
    for kind in ("plain", "filtered"):
        eng.clear_masks()
        res = eng.profile_search_batch(profs, masks) if kind == "plain" else eng.profile_search_batch_filtered(profs, masks)
        h = hashlib.sha1()
        for r in res:
            h.update(r.bytes)
        out[kind + "_sha1"] = h.hexdigest()
    eng.close()
    line = json.dumps(out)
    print(line)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
