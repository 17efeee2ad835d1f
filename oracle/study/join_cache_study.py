#!/usr/bin/env python
"""Design study (test infrastructure, CPU only): how much of the join's all-pairs scan is repeated work
across the gap-size schedule?  Runs the oracle built with -DORACLE_STUDY (make -C oracle study) on a
sample of list pairs of a BASELINE-shaped workload and prints the counters of oracle/study/join_cache.inc.

    python oracle/study/join_cache_study.py [--workload c4|c3|d2] [--pairs 400]
"""
import argparse
import ctypes as C
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python"))
from iadhore_b200 import capi, ingest, synth  # noqa: E402

LIB = os.path.join(HERE, "liboracle_study.so")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--pairs", type=int, default=400)
    ap.add_argument("--j2", action="store_true", help="run the lane-level model of KT<32>::join_pass2 against the reference's merges")
    ap.add_argument("--spec", type=int, default=0, help="speculative merge batches of up to K disjoint pairs (0 = off)")
    a = ap.parse_args()
    if a.workload == "c4":      # config 4's per-GHM shape (2,500-gene lists, 20,000 families), fewer genomes
        cfg = synth.config_c4(1.0)
        cfg.genomes = cfg.genomes[:6]
    elif a.workload == "c3":
        cfg = synth.config_c3(1.0)
        cfg.genomes = cfg.genomes[:5]
    else:
        cfg = synth.config_d2(0.25)
    ds = ingest.dataset_from_synth(cfg)
    eng = capi.Engine(LIB, "oracle_", ds)
    eng.lib.oracle_study_set_spec(C.c_int(a.spec))
    eng.lib.oracle_study_set_j2_model(C.c_int(1 if a.j2 else 0))
    eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    n = ds.n_lists
    allp = [(x, y) for x in range(n) for y in range(x + 1)]
    step = max(1, len(allp) // a.pairs)
    sel = allp[::step][:a.pairs]
    r = eng.level2_pairs([p[0] for p in sel], [p[1] for p in sel])
    out = (C.c_longlong * 33)()
    eng.lib.oracle_study_counters(out)
    ghms, calls, pairs, near_gap, near_max, first_max, clusters, changed = list(out)[:8]
    merges, max_clusters = out[8], out[9]
    labels = ['0', '1', '2-4', '5-16', '17-64', '>64']
    print(json.dumps(dict(workload=a.workload, list_pairs=len(sel), dots=r.stats["n_dots"], multiplicons=r.n_multiplicons,
                          ghms_clustered=ghms, join_calls=calls, clusters_per_call=clusters / max(calls, 1),
                          changed_per_call=changed / max(calls, 1), pairs_scanned=pairs, near_at_gap=near_gap,
                          near_at_max_gap=near_max, first_time_at_max_gap=first_max,
                          repeat_factor_today=near_gap / max(first_max, 1),
                          scan_factor=pairs / max(first_max, 1), merges=merges, max_clusters=max_clusters,
                          join_pass2_model=dict(merges_checked=out[31], mismatches=out[32]) if a.j2 else None,
                          warp_view=dict(row_chunks_today=out[27], chunks_running_the_evaluation_today=out[28],
                                         evaluation_steps_two_phase=out[29], scan_steps_two_phase=out[30]),
                          spec=dict(k=a.spec, batches=out[22], offered=out[23], accepted=out[24],
                                    merges_per_batch=merges / max(out[22], 1),
                                    criterion_unsafe_batches=out[25], criterion_lost_merges=out[26]) if a.spec else None,
                          calls_by_clusters={l: out[10 + i] for i, l in enumerate(labels)},
                          pairs_by_clusters={l: out[16 + i] for i, l in enumerate(labels)})))


if __name__ == "__main__":
    main()
