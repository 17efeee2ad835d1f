#!/usr/bin/env python
"""GPU fuzz: the CUDA path against the oracle on random small datasets with NON-DEFAULT settings
(the same generator the CPU test tests/test_oracle_fuzz.py pins the oracle with against the reference).
Level-2 call over all pairs, dot sets, then profile searches with accumulating masks.

    python tools/fuzz_parity.py [--seed 1] [--cases 40]      # on a B200; exit code = number of failing cases

Written at the end of round 1, after the GPU budget was spent: the device path has so far only been
compared under the shipped default settings (all three multiple-hypothesis modes, both table types).
Run this first in the next round; every failure is a parity bug to fix before any tuning.
"""
import argparse
import json
import os
import random
import sys

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python"))
sys.path.insert(0, os.path.join(REPO, "tests"))
from iadhore_b200 import capi, fuzzcfg, ingest  # noqa: E402
import parity_lib as pl  # noqa: E402


def one(cfg):
    ds = ingest.dataset_from_synth(cfg)
    gpu, cpu = capi.load_product(ds), pl.oracle_engine(ds)
    errs = []
    try:
        for e in (gpu, cpu):
            e.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, fuzzcfg.MH[cfg.mult_hyp], cfg.q_value, cfg.prob_cutoff)
        a, b = gpu.level2_all_pairs(), cpu.level2_all_pairs()
        msg = capi.compare_packed(a.bytes, b.bytes, pl.P_RTOL)
        if msg:
            errs.append("L2: " + msg)
        n = ds.n_lists
        xs = [x for x in range(n) for y in range(x + 1)]
        ys = [y for x in range(n) for y in range(x + 1)]
        for (u, v), x, y in zip(zip(gpu.build_dots(xs, ys), cpu.build_dots(xs, ys)), xs, ys):
            if not (np.array_equal(u[0], v[0]) and np.array_equal(u[1], v[1])):
                errs.append(f"dots ({x},{y})")
        ms = capi.parse_packed(b.bytes)
        order = sorted(range(len(ms)), key=lambda i: -sum(len(c.aps) for c in ms[i].clusters))[:10]
        for i in order:
            m = ms[i]
            prof = pl.simple_profile(ds, m.x_object_id, m.y_object_id, m)
            for e in (gpu, cpu):
                e.mask_range(m.x_object_id, m.bounds[0], m.bounds[1])
                e.mask_range(m.y_object_id, m.bounds[2], m.bounds[3])
            msg = capi.compare_packed(gpu.profile_search(prof).bytes, cpu.profile_search(prof).bytes, pl.P_RTOL)
            if msg:
                errs.append(f"HL profile from multiplicon {i}: " + msg)
        return errs, a.n_multiplicons
    finally:
        gpu.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cases", type=int, default=40)
    a = ap.parse_args()
    rng = random.Random(a.seed)
    bad = 0
    for k in range(a.cases):
        cfg = fuzzcfg.random_config(rng)
        errs, nm = one(cfg)
        print(("OK  " if not errs else "FAIL"), json.dumps(fuzzcfg.describe(cfg)), nm, errs[:3], flush=True)
        bad += bool(errs)
    print(f"{bad} failing of {a.cases}")
    sys.exit(min(bad, 100))


if __name__ == "__main__":
    main()
