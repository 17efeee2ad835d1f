#!/usr/bin/env python
"""Where does the end-to-end (fresh context per step) time go?  Wall time of every C-ABI call."""
import os, sys, time
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python")); sys.path.insert(0, REPO)
from iadhore_b200 import capi, ingest
import bench
cfg = bench.workload_config("c3")
ds = ingest.dataset_from_synth(cfg)
eng = capi.load_product(ds)
eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
r0 = eng.level2_all_pairs()
profs = bench.make_profiles(ds, r0.bytes, 32)
eng.close()
for step in range(3):
    t = [time.perf_counter()]
    e2 = capi.load_product(ds); t.append(time.perf_counter())
    e2.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    e2.clear_masks()
    r = e2.level2_all_pairs(); t.append(time.perf_counter())
    hl = []
    for m, p in profs:
        a = time.perf_counter()
        e2.mask_range(m.x_object_id, m.bounds[0], m.bounds[1]); e2.mask_range(m.y_object_id, m.bounds[2], m.bounds[3])
        pr = e2.profile_search(p)
        hl.append((time.perf_counter() - a, pr.stats["total_ms"]))
    t.append(time.perf_counter())
    e2.close(); t.append(time.perf_counter())
    print(f"step {step}: create {1e3*(t[1]-t[0]):.1f} L2 {1e3*(t[2]-t[1]):.1f} (dev {r.stats['total_ms']:.1f}) HL {1e3*(t[3]-t[2]):.1f} "
          f"(sum dev {sum(h[1] for h in hl):.1f}; first3 wall {[round(1e3*h[0],2) for h in hl[:3]]}) close {1e3*(t[4]-t[3]):.1f} ms")
