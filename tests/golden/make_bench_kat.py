"""Known answers for the profile batch bench.py times on config 4: the reference's own first 32 aligned profiles
(bench_profiles_c4.bin.gz, with the Link sets their searches inherit: bench_profile_links_c4.bin.xz) searched by the
ORACLE against all 500 lists, without and with the homolog-count filter -> tests/golden/bench_profiles_c4.kat.json
(digest of the 32 packed streams, parity_lib.stream_digest), plus the digest of the whole level-2 call of a bench step
(125,250 list pairs) and of config 3's (5,050).  The GPU suite compares the device's batch with it
(tests/test_gpu_bench_parity.py, tests/test_workloads_full_size.py).  CPU only, about 10 minutes on 8 cores.

    python tests/golden/make_bench_kat.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python"))
sys.path.insert(0, REPO)
import parity_lib as pl  # noqa: E402
from iadhore_b200 import capi, ingest, profiles  # noqa: E402
import bench  # noqa: E402


def load(ds, workload="c4", n=32):
    path = bench.ref_profiles_path(workload)
    return profiles.load_reference_profiles(ds, path, n, links_path=path.replace("bench_profiles_", "bench_profile_links_").replace(".bin.gz", ".bin.xz"))


def level2_digest(workload):
    """The whole level-2 call of a workload (every list pair in the reference's pair order, one stream)."""
    cfg = bench.workload_config(workload)
    ds = ingest.dataset_from_synth(cfg)
    eng = pl.oracle_engine(ds)
    eng.lib.oracle_set_threads(eng.ctx, os.cpu_count() or 1)
    eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    l2 = eng.level2_all_pairs()
    d = pl.stream_digest([l2.bytes])
    d["ghms"] = int(l2.stats["n_ghm"])
    return d


def main():
    path = os.path.join(HERE, "bench_profiles_c4.kat.json")
    if sys.argv[1:] == ["c3"]:   # only (re)make the config-3 entry
        out = json.load(open(path))
        out["c3_level2"] = level2_digest("c3")
        with open(path, "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out["c3_level2"]))
        return
    cfg = bench.workload_config("c4")
    ds = ingest.dataset_from_synth(cfg)
    bp = load(ds)
    eng = pl.oracle_engine(ds)
    eng.lib.oracle_set_threads(eng.ctx, os.cpu_count() or 1)
    eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    plain, filtered = [], []
    for k, (masks, p) in enumerate(bp):
        for (l, b, e) in masks:
            eng.mask_range(l, b, e)
        plain.append(eng.profile_search(p).bytes)
        filtered.append(eng.profile_search_filtered(p).bytes)
        print(k, len(plain[-1]), len(filtered[-1]), flush=True)
    out = {"workload": "c4", "profiles": len(bp), "plain": pl.stream_digest(plain), "filtered": pl.stream_digest(filtered)}
    # the level-2 call of a bench step: all 125,250 list pairs in the reference's pair order, one stream
    eng.clear_masks()
    l2 = eng.level2_all_pairs()
    out["level2"] = pl.stream_digest([l2.bytes])
    out["level2"]["ghms"] = int(l2.stats["n_ghm"])
    out["c3_level2"] = level2_digest("c3")   # config 3 (5,050 list pairs), an extra configuration of the bench line
    with open(os.path.join(HERE, "bench_profiles_c4.kat.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
