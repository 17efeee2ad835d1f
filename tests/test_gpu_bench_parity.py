"""Parity of the CUDA path ON THE BENCHED CONFIGURATIONS ("next round" item 1).

bench.py quotes its number on BASELINE config 4 (125,250 list pairs in one level-2 call, the
throughput path of the clustering kernels) plus profile searches; these tests put exactly those
calls under the oracle:
  * the FULL config-4 level-2 call is run once, as bench.py runs it; the packed blocks of a sample
    of >= 3,000 GHMs (the 500 heaviest + 2,500 random ones) are cut out of its stream and compared
    with the oracle's bytes for the same list pairs; the same sample is also run as a call of its own
    on the fused path and with the throughput path forced (IADH_PASS_MIN=1 IADH_PASS_RATIO=0);
  * the digest of the WHOLE config-4 level-2 stream against the known answer computed by the oracle on the CPU
    (tests/golden/bench_profiles_c4.kat.json; the benched profile batch and the datasetII-shaped workload follow in
    tests/test_workloads_full_size.py);
  * config 3: the full level-2 call against the oracle on a sample, then 20 profile searches with
    accumulating masks against the oracle;
  * 100 random small datasets with NON-default settings (tools/fuzz_parity.py as a test).
Bar: bytes identical, p-values within parity_lib.P_RTOL (1e-12 relative, north_star).
The oracle port is multi-threaded here (it is the checker, nothing is timed)."""
import ctypes as C
import os
import random
import struct
import sys

import numpy as np
import pytest

import parity_lib as pl
from iadhore_b200 import capi, fuzzcfg, ingest, synth

sys.path.insert(0, os.path.join(pl.REPO, "tools"))

pytestmark = pytest.mark.gpu


def split_packed(b: bytes):
    """Byte offsets of the multiplicons of a packed stream (Multiplicon.cpp:210-264): off[i]..off[i+1]."""
    (n,) = struct.unpack_from("<i", b, 0)
    off = 4
    offs = np.empty(n + 1, np.int64)
    u32 = struct.Struct("<i").unpack_from
    for i in range(n):
        offs[i] = off
        nc = u32(b, off + 29)[0]
        off += 33
        for _ in range(nc):
            off += 102 + 17 * u32(b, off + 98)[0]
    offs[n] = off
    assert off == len(b)
    return offs


def blocks_of(res, ghms):
    """Concatenated packed multiplicons of the given GHM indices (in that order) of a result."""
    b = res.bytes
    offs = split_packed(b)
    gi = np.asarray(res.ghm_index)
    first = np.searchsorted(gi, ghms, side="left")
    last = np.searchsorted(gi, ghms, side="right")
    parts, n = [], 0
    for a, z in zip(first, last):
        parts.append(b[offs[a]:offs[z]])
        n += int(z - a)
    return struct.pack("<i", n) + b"".join(parts)


def sample_of(res, n_pairs, n_heavy, n_random, seed):
    """The n_heavy GHMs with the most multiplicons in the full result + n_random others."""
    per = np.bincount(np.asarray(res.ghm_index), minlength=n_pairs)
    heavy = np.argsort(-per, kind="stable")[:n_heavy]
    rest = np.setdiff1d(np.arange(n_pairs), heavy)
    rnd = np.random.RandomState(seed).choice(rest, min(n_random, len(rest)), replace=False)
    return np.sort(np.concatenate([heavy, rnd]))


def _engines(cfg):
    ds = ingest.dataset_from_synth(cfg)
    gpu, cpu = capi.load_product(ds), pl.oracle_engine(ds)
    cpu.lib.oracle_set_threads(cpu.ctx, C.c_int(min(32, os.cpu_count() or 1)))
    for e in (gpu, cpu):
        e.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    return ds, gpu, cpu


@pytest.fixture(scope="module")
def c4():
    cfg = synth.config_c4(1.0)
    ds, gpu, cpu = _engines(cfg)
    xs, ys = capi.plan_pairs(np.diff(ds.list_off))
    full = gpu.level2_all_pairs()                      # the benched call, on the path bench.py takes
    sel = sample_of(full, len(xs), 500, 2500, seed=4)
    want = cpu.level2_pairs(xs[sel], ys[sel])
    yield dict(cfg=cfg, ds=ds, gpu=gpu, xs=xs, ys=ys, full=full, sel=sel, want=want)
    gpu.close()


def test_c4_full_call_matches_oracle_on_sample(c4):
    full, sel, want = c4["full"], c4["sel"], c4["want"]
    assert full.stats["n_ghm"] == 125250 and len(sel) >= 3000
    assert want.n_multiplicons > 3000
    got = blocks_of(full, sel)
    assert capi.compare_packed(got, want.bytes, pl.P_RTOL) is None


@pytest.mark.parametrize("path", ["default", "throughput"])
def test_c4_sample_call_matches_oracle(c4, path, monkeypatch):
    """The sample as a call of its own: `default` lets the engine choose (3,000 GHMs: latency path),
    `throughput` forces the path of large batches on it."""
    if path == "throughput":
        monkeypatch.setenv("IADH_PASS_MIN", "1")
        monkeypatch.setenv("IADH_PASS_RATIO", "0")
    cfg = c4["cfg"]
    gpu = capi.load_product(c4["ds"])
    gpu.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    r = gpu.level2_pairs(c4["xs"][c4["sel"]], c4["ys"][c4["sel"]])
    gpu.close()
    assert capi.compare_packed(r.bytes, c4["want"].bytes, pl.P_RTOL) is None
    assert np.array_equal(r.ghm_index, c4["want"].ghm_index)


def test_c4_profile_searches_match_oracle(c4):
    """The profile searches of a bench step (same construction as bench.make_profiles), first 6 of 32,
    masks accumulating, against the oracle."""
    import bench
    cfg, ds, gpu = c4["cfg"], c4["ds"], c4["gpu"]
    cpu = pl.oracle_engine(ds)
    cpu.lib.oracle_set_threads(cpu.ctx, C.c_int(min(32, os.cpu_count() or 1)))
    cpu.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    profs = bench.make_profiles(ds, c4["full"].bytes, 6)
    gpu.clear_masks()
    found = 0
    for m, p in profs:
        for e in (gpu, cpu):
            e.mask_range(m.x_object_id, m.bounds[0], m.bounds[1])
            e.mask_range(m.y_object_id, m.bounds[2], m.bounds[3])
        a, b = gpu.profile_search(p), cpu.profile_search(p)
        assert capi.compare_packed(a.bytes, b.bytes, pl.P_RTOL) is None
        found += a.n_multiplicons
    gpu.clear_masks()
    assert found > 0


def test_c4_full_level2_stream_matches_the_oracle_known_answer(c4):
    """The WHOLE level-2 call of a bench step (125,250 list pairs, one stream of ~320 MB) against the oracle's digest of
    the same call (tests/golden/bench_profiles_c4.kat.json, "level2": made on the CPU by tests/golden/make_bench_kat.py --
    minutes of oracle time, so the answer is committed): every multiplicon of the benched call, not a sample."""
    import json
    kat = json.load(open(os.path.join(pl.GOLDEN, "bench_profiles_c4.kat.json")))["level2"]
    full = c4["full"]
    assert full.stats["n_ghm"] == kat["ghms"] == 125250
    msg = pl.digest_mismatch(pl.stream_digest([full.bytes]), kat)
    assert msg is None, msg


def test_c3_level2_and_20_profile_searches_match_oracle():
    cfg = synth.config_c3(1.0)
    ds, gpu, cpu = _engines(cfg)
    xs, ys = capi.plan_pairs(np.diff(ds.list_off))
    full = gpu.level2_all_pairs()
    assert full.stats["n_ghm"] == 5050
    sel = sample_of(full, len(xs), 500, 2500, seed=3)
    want = cpu.level2_pairs(xs[sel], ys[sel])
    assert capi.compare_packed(blocks_of(full, sel), want.bytes, pl.P_RTOL) is None
    # ... and the whole stream (8,053 multiplicons) against the oracle's committed digest of the same call
    import json
    kat = json.load(open(os.path.join(pl.GOLDEN, "bench_profiles_c4.kat.json")))["c3_level2"]
    msg = pl.digest_mismatch(pl.stream_digest([full.bytes]), kat)
    assert msg is None, msg
    import bench
    profs = bench.make_profiles(ds, full.bytes, 20)
    found = 0
    for m, p in profs:
        for e in (gpu, cpu):
            e.mask_range(m.x_object_id, m.bounds[0], m.bounds[1])
            e.mask_range(m.y_object_id, m.bounds[2], m.bounds[3])
        a, b = gpu.profile_search(p), cpu.profile_search(p)
        assert capi.compare_packed(a.bytes, b.bytes, pl.P_RTOL) is None
        found += a.n_multiplicons
    assert found > 0
    gpu.close()


def test_fuzz_100_non_default_settings():
    """tools/fuzz_parity.py --seed 1 --cases 100 as a test: gaps 3-60, cluster gaps, q 0.3-0.95,
    cut-offs 0.001-1, 3-7 anchor points, all filter modes, both table types."""
    import fuzz_parity
    rng = random.Random(1)
    bad = []
    for k in range(100):
        cfg = fuzzcfg.random_config(rng)
        errs, _ = fuzz_parity.one(cfg)
        if errs:
            bad.append((fuzzcfg.describe(cfg), errs[:2]))
    assert not bad, bad[:3]
