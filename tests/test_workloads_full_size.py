"""Full-size parity of the remaining calls of the bench line, written at the very end of round 2 (after the round's GPU
minutes were spent; prepared on the CPU, see profiles/parity_r02.md) -- kept in a file of their own, collected last:

  * the profile batch bench.py times on config 4 (the reference's own first 32 aligned profiles x 500 lists in one call),
    without and with the homolog-count filter, against the oracle's committed digests;
  * the datasetII-shaped workload (`--workload d2`): the whole level-2 call and 8 profile searches against the oracle, live.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import parity_lib as pl
from iadhore_b200 import capi, ingest, synth

sys.path.insert(0, os.path.join(pl.REPO, "tools"))

pytestmark = pytest.mark.gpu


def _engines(cfg, oracle=True):
    ds = ingest.dataset_from_synth(cfg)
    gpu = capi.load_product(ds)
    cpu = pl.oracle_engine(ds) if oracle else None
    if cpu is not None:
        cpu.lib.oracle_set_threads(cpu.ctx, C.c_int(min(32, os.cpu_count() or 1)))
    for e in (gpu, cpu):
        if e is not None:
            e.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    return ds, gpu, cpu


@pytest.fixture(scope="module")
def c4():
    ds, gpu, _ = _engines(synth.config_c4(1.0), oracle=False)
    yield dict(ds=ds, gpu=gpu)
    gpu.close()


def test_c4_benched_profile_batch_matches_the_oracle_known_answers(c4):
    """The profile batch bench.py times -- the reference's own first 32 aligned profiles of config 4 against all 500
    lists in one iadh_profile_search_batch call -- against the oracle's digest of the same 32 searches
    (tests/golden/bench_profiles_c4.kat.json, made on the CPU by tests/golden/make_bench_kat.py; the oracle needs
    minutes for them, so the answers are committed instead of recomputed here).  Then the same batch through
    iadh_profile_search_batch_filtered (homolog-count filter on the device, the Link sets of the 32 profiles)."""
    import json
    sys.path.insert(0, os.path.join(pl.REPO, "tests", "golden"))
    import make_bench_kat
    kat = json.load(open(os.path.join(pl.GOLDEN, "bench_profiles_c4.kat.json")))
    ds, gpu = c4["ds"], c4["gpu"]
    bp = make_bench_kat.load(ds)
    assert len(bp) == kat["profiles"]
    profs, masks = [p for _, p in bp], [m for m, _ in bp]
    for kind in ("plain", "filtered"):
        gpu.clear_masks()
        res = gpu.profile_search_batch(profs, masks) if kind == "plain" else gpu.profile_search_batch_filtered(profs, masks)
        msg = pl.digest_mismatch(pl.stream_digest([r.bytes for r in res]), kat[kind])
        assert msg is None, (kind, msg)
        assert sum(r.stats["n_ghm"] for r in res) == 32 * 500
    gpu.clear_masks()
    assert kat["filtered"]["multiplicons"] < kat["plain"]["multiplicons"]


def test_d2_level2_and_profile_searches_match_oracle():
    """The datasetII-shaped workload of the bench line (`--workload d2`: 420 ragged lists, 7 ... 9,987 genes, the shipped
    list-length distribution with synthetic genes): the WHOLE level-2 call (88,410 list pairs, GHMs up to 6,000 dots)
    byte for byte against the oracle, then 8 profile searches with accumulating masks one call at a time and the same 8
    as one batch."""
    import bench
    cfg = bench.workload_config("d2")
    ds, gpu, cpu = _engines(cfg)
    full, want = gpu.level2_all_pairs(), cpu.level2_all_pairs()
    assert full.stats["n_ghm"] == want.stats["n_ghm"] == 88410 and full.stats["n_dots"] == want.stats["n_dots"]
    assert capi.compare_packed(full.bytes, want.bytes, pl.P_RTOL) is None
    assert np.array_equal(full.ghm_index, want.ghm_index)
    profs = bench.make_profiles(ds, want.bytes, 8)
    masks = [[(m.x_object_id, m.bounds[0], m.bounds[1]), (m.y_object_id, m.bounds[2], m.bounds[3])] for m, _ in profs]
    single, found = [], 0
    for (m, p), mk in zip(profs, masks):
        for e in (gpu, cpu):
            for (l, b, e2) in mk:
                e.mask_range(l, b, e2)
        a, b = gpu.profile_search(p), cpu.profile_search(p)
        assert capi.compare_packed(a.bytes, b.bytes, pl.P_RTOL) is None
        single.append(a.bytes)
        found += a.n_multiplicons
    assert found > 50
    gpu.clear_masks()
    for k, r in enumerate(gpu.profile_search_batch([p for _, p in profs], masks)):
        assert r.bytes == single[k], k
    gpu.close()
