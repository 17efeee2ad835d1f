#!/usr/bin/env python
"""K1 (GHM build) alone on a dense workload: achieved algorithmic GB/s (5(|X|+|Y|)+4D per pair).

    python tools/build_probe.py [--workload c5|c3] [--scale 1.0] [--reps 5]
"""
import argparse, os, sys, time, json
import numpy as np
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python")); sys.path.insert(0, REPO)
from iadhore_b200 import capi, ingest, synth

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c5")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
cfg = synth.config_c5(a.scale) if a.workload == "c5" else synth.config_c3(a.scale)
t0 = time.time()
ds = ingest.dataset_from_synth(cfg)
print(f"{cfg.name}: {ds.n_lists} lists, {len(ds.fam)} remapped genes, setup {time.time()-t0:.1f}s", file=sys.stderr)
sizes = np.diff(ds.list_off)
xs, ys = capi.plan_pairs(sizes)
eng = capi.load_product(ds)
for r in range(a.reps):
    st = eng.build_only(xs, ys)
    alg = 5 * int(sizes[xs].sum() + sizes[ys].sum()) + 4 * st["n_dots"]
    print(json.dumps({"pairs": len(xs), "dots": st["n_dots"], "build_ms": st["build_ms"], "fill_ms": st["cluster_ms"],
                      "alg_bytes": alg, "GBps": alg / st["build_ms"] / 1e6, "dots_per_s": st["n_dots"] / st["build_ms"] * 1e3}))
eng.close()
