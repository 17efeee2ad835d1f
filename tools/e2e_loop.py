#!/usr/bin/env python
"""The e2e loop of bench.py alone (fresh context from host buffers per step, config 4): wall time per step and, with
IADH_TRACE_ALLOC=1, every block the library had to get from the driver instead of its cache."""
import os, sys, time
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python")); sys.path.insert(0, REPO)
from iadhore_b200 import capi, ingest
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
cfg = bench.workload_config(wl)
ds = ingest.dataset_from_synth(cfg, device=0)
eng = capi.load_product(ds)
eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
profs, _ = bench.bench_profiles(ds, wl, 32, lambda: eng.level2_all_pairs().bytes)
for _ in range(4):
    bench.hot_path_step(eng, profs, 0, 1)
# as bench.py does after the timed loop: single searches, then the resident context is closed
eng.clear_masks()
for masks, p in profs[:8]:
    for (l, b, e) in masks:
        eng.mask_range(l, b, e)
    eng.profile_search(p)
eng.close()
if os.environ.get("E2E_EXTRAS"):   # as bench.py: short runs of two other workloads in between
    for xwl in ("c3", "d2"):
        xcfg = bench.workload_config(xwl)
        xds = ingest.dataset_from_synth(xcfg, device=0)
        xe = capi.load_product(xds)
        xe.set_params(xcfg.gap_size, xcfg.cluster_gap, xcfg.anchor_points, capi.MH_BONFERRONI, xcfg.q_value, xcfg.prob_cutoff)
        xp, _ = bench.bench_profiles(xds, xwl, 32, lambda: xe.level2_all_pairs().bytes)
        for _ in range(4):
            bench.hot_path_step(xe, xp, 0, 1)
        xe.close()
    print("---- extras done", file=sys.stderr, flush=True)
for step in range(6):
    print(f"---- step {step}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    e2 = capi.load_product(ds)
    t1 = time.perf_counter()
    e2.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    bench.hot_path_step(e2, profs, 0, 1)
    t2 = time.perf_counter()
    e2.close()
    t3 = time.perf_counter()
    print(f"step {step}: create {1e3 * (t1 - t0):.1f} ms, step {1e3 * (t2 - t1):.1f} ms, close {1e3 * (t3 - t2):.1f} ms, total {1e3 * (t3 - t0):.1f} ms", file=sys.stderr, flush=True)
