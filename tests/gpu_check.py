"""Standalone GPU parity run (same checks as `pytest -m gpu`, with timing prints).

    python tests/gpu_check.py [--size small|medium|c3s0.3] [--skip-golden]
"""
import argparse
import sys
import time

import numpy as np

import parity_lib as pl
from iadhore_b200 import capi, ingest, synth
from trace_io import read_trace


def check_golden():
    bad = 0
    for path in pl.golden_traces():
        tr = read_trace(path)
        eng = capi.load_product(tr.dataset)
        t0 = time.time()
        errs = pl.replay_trace(eng, tr, pl.P_RTOL)
        print(f"{path.split('/')[-1]:24s} {'OK' if not errs else 'FAIL'} ({time.time() - t0:.2f}s)")
        for e in errs[:8]:
            print("   ", e)
        bad += len(errs)
        eng.close()
    return bad


def check_synth(cfg, n_profiles=40, mult_hyp=capi.MH_BONFERRONI):
    ds = ingest.dataset_from_synth(cfg)
    print(f"{cfg.name}: {ds.n_lists} lists, {len(ds.fam)} remapped genes")
    gpu = capi.load_product(ds)
    cpu = pl.oracle_engine(ds)
    for e in (gpu, cpu):
        e.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, mult_hyp, cfg.q_value, cfg.prob_cutoff)
    t0 = time.time()
    rg = gpu.level2_all_pairs()
    t1 = time.time()
    rc = cpu.level2_all_pairs()
    t2 = time.time()
    msg = capi.compare_packed(rg.bytes, rc.bytes, pl.P_RTOL)
    exact = rg.bytes == rc.bytes
    print(f"  L2: {rg.stats['n_ghm']} GHMs, {rg.stats['n_dots']} dots, {rg.n_multiplicons} multiplicons; "
          f"gpu {t1 - t0:.3f}s (build {rg.stats['build_ms']:.2f} ms, cluster {rg.stats['cluster_ms']:.2f} ms), "
          f"oracle {t2 - t1:.3f}s; parity={'OK' if not msg else msg} bit-exact-incl-p={exact}")
    bad = 1 if msg else 0
    # profile searches on 2-segment profiles made from the largest L2 multiplicons
    ms = capi.parse_packed(rc.bytes)
    ordered = sorted(range(len(ms)), key=lambda i: -sum(len(c.aps) for c in ms[i].clusters))[:n_profiles]
    tg = tc = 0.0
    nb = 0
    for i in ordered:
        m = ms[i]
        prof = pl.simple_profile(ds, m.x_object_id, m.y_object_id, m)
        for e in (gpu, cpu):
            e.mask_range(m.x_object_id, m.bounds[0], m.bounds[1])
            e.mask_range(m.y_object_id, m.bounds[2], m.bounds[3])
        a = time.time()
        pg = gpu.profile_search(prof)
        b = time.time()
        pc = cpu.profile_search(prof)
        c = time.time()
        tg += b - a
        tc += c - b
        mm = capi.compare_packed(pg.bytes, pc.bytes, pl.P_RTOL)
        if mm:
            nb += 1
            if nb < 5:
                print(f"  HL profile from multiplicon {i}: {mm}")
    print(f"  HL: {len(ordered)} profiles x {ds.n_lists} lists: gpu {tg:.3f}s oracle {tc:.3f}s mismatches {nb}")
    gpu.close()
    cpu.close()
    return bad + nb


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="small")
    ap.add_argument("--skip-golden", action="store_true")
    ap.add_argument("--trace", default=None, help="replay one (large) reference trace")
    a = ap.parse_args()
    bad = 0
    if a.trace:
        tr = read_trace(a.trace)
        eng = capi.load_product(tr.dataset)
        t0 = time.time()
        errs = pl.replay_trace(eng, tr, pl.P_RTOL)
        print(f"{a.trace}: {'OK' if not errs else 'FAIL'} ({time.time() - t0:.2f}s), "
              f"{sum(1 for e in tr.events if e[0] == 'l2')} L2 + {sum(1 for e in tr.events if e[0] == 'hl')} HL GHMs")
        for e in errs[:10]:
            print("   ", e)
        errs0 = pl.replay_trace(eng, tr, 0.0, check_dots=False)
        print("   p-values bit-exact too:", not errs0)
        bad += len(errs)
    if not a.skip_golden:
        bad += check_golden()
    if a.size == "small":
        bad += check_synth(synth.config_small())
        bad += check_synth(synth.config_small(seed=22), mult_hyp=capi.MH_FDR)
    elif a.size == "medium":
        bad += check_synth(synth.config_c3(0.2))
    elif a.size.startswith("c3s"):
        bad += check_synth(synth.config_c3(float(a.size[3:])))
    print("TOTAL MISMATCHES", bad)
    sys.exit(1 if bad else 0)
