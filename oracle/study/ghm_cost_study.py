#!/usr/bin/env python
"""Design study (test infrastructure, CPU only): per-GHM work of level-2 GHMs and of the GHMProfiles of
profile searches, on a BASELINE-shaped workload -- sum (throughput) against maximum (the latency of one
profile search is its heaviest GHM).  Uses the oracle built with -DORACLE_STUDY (make -C oracle study).

    python oracle/study/ghm_cost_study.py [--workload c4|c3] [--pairs 300] [--profiles 16]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python"))
from iadhore_b200 import capi, ingest, synth  # noqa: E402
from iadhore_b200.profiles import simple_profile  # noqa: E402

LIB = os.path.join(HERE, "liboracle_study.so")
COLS = ["dots", "pairs_scanned", "near_pairs", "merges", "enrich_tests", "max_clusters", "level", "batches"]


def records(eng):
    cap = 1 << 20
    buf = (C.c_longlong * (8 * cap))()
    eng.lib.oracle_study_records.restype = C.c_longlong
    n = eng.lib.oracle_study_records(buf, C.c_longlong(cap))
    return np.ctypeslib.as_array(buf)[:8 * min(n, cap)].reshape(-1, 8).copy()


def summary(r):
    out = {"ghms": int(len(r))}
    for i, c in enumerate(COLS):
        if c == 'level':
            continue
        v = r[:, i]
        out[c] = dict(sum=int(v.sum()), mean=float(v.mean()), p50=float(np.percentile(v, 50)), p99=float(np.percentile(v, 99)),
                      max=int(v.max()))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--pairs", type=int, default=300)
    ap.add_argument("--profiles", type=int, default=16)
    ap.add_argument("--j2", action="store_true", help="run the lane-level model of KT<32>::join_pass2 against the reference's merges")
    ap.add_argument("--spec", type=int, default=0, help="speculative merge batches of up to K disjoint pairs (0 = off)")
    a = ap.parse_args()
    cfg = synth.config_c4(1.0) if a.workload == "c4" else synth.config_c3(1.0)
    cfg.genomes = cfg.genomes[:6]
    ds = ingest.dataset_from_synth(cfg)
    eng = capi.Engine(LIB, "oracle_", ds)
    eng.lib.oracle_study_set_spec(C.c_int(a.spec))
    eng.lib.oracle_study_set_j2_model(C.c_int(1 if a.j2 else 0))
    eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    n = ds.n_lists
    allp = [(x, y) for x in range(n) for y in range(x + 1)]
    sel = allp[::max(1, len(allp) // a.pairs)][:a.pairs]
    r = eng.level2_pairs([p[0] for p in sel], [p[1] for p in sel])
    out = {"workload": a.workload, "level2": summary(records(eng))}
    # profile searches as bench.py makes them: the largest level-2 multiplicons, masks accumulating
    ms = capi.parse_packed(r.bytes)
    order = sorted(range(len(ms)), key=lambda i: -sum(len(c.aps) for c in ms[i].clusters))[:a.profiles]
    per_search = []
    allr = []
    for i in order:
        m = ms[i]
        prof = simple_profile(ds, m.x_object_id, m.y_object_id, m)
        eng.mask_range(m.x_object_id, m.bounds[0], m.bounds[1])
        eng.mask_range(m.y_object_id, m.bounds[2], m.bounds[3])
        eng.profile_search(prof)
        rr = records(eng)
        allr.append(rr)
        w = rr[:, 1] + 50 * rr[:, 3] + rr[:, 4] / 8   # crude work proxy: pair scans + 50 per merge + enrich tests / 8
        per_search.append(dict(L=int(prof.L), ghms=int(len(rr)), dots_sum=int(rr[:, 0].sum()), dots_max=int(rr[:, 0].max()),
                               heaviest=dict(zip(COLS, map(int, rr[int(w.argmax())]))),
                               heaviest_share_of_work=float(w.max() / max(w.sum(), 1))))
    out["profile_search"] = summary(np.concatenate(allr))
    out["per_search"] = per_search
    if a.spec:
        cnt = (C.c_longlong * 33)()
        eng.lib.oracle_study_counters(cnt)
        out["speculation"] = dict(k=a.spec, batches=cnt[22], offered=cnt[23], accepted=cnt[24],
                                  criterion_unsafe_batches=cnt[25], criterion_lost_merges=cnt[26])
    if a.j2:
        cnt = (C.c_longlong * 33)()
        eng.lib.oracle_study_counters(cnt)
        out["join_pass2_model"] = dict(merges_checked=cnt[31], mismatches=cnt[32])
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
