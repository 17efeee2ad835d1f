"""Parity of the CUDA path (through the C-ABI, libiadhore_b200.so) against
 (1) golden traces of the unmodified reference, (2) the oracle on seeded synthetic genomes,
 (3) size-independent properties at the full C3 shape.
Bar (north_star): anchor points, cluster membership, bounds, levels, regression doubles
bit-exact; p-values within 1e-12 relative (written in parity_lib.P_RTOL)."""
import os

import numpy as np
import pytest

import parity_lib as pl
from iadhore_b200 import capi, ingest, synth
from trace_io import read_trace

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k1", ["default", "tiles", "expansion"])
@pytest.mark.parametrize("path", pl.golden_traces(), ids=lambda p: p.split("/")[-1].split(".")[0])
def test_cuda_replays_reference_trace(path, k1, monkeypatch):
    # K1 has two build paths (tile kernels / family expansion, chosen by batch size): the traces replay on both
    if k1 != "default":
        monkeypatch.setenv("IADH_EXPAND_MIN", "1" if k1 == "expansion" else "1000000000")
        monkeypatch.setenv("IADH_HL_EXPAND_MIN", "1" if k1 == "expansion" else "1000000000")
    tr = read_trace(path)
    eng = capi.load_product(tr.dataset)
    errs = pl.replay_trace(eng, tr, rtol_p=pl.P_RTOL)
    eng.close()
    assert not errs, errs[:5]


def _pair(cfg, mult_hyp):
    ds = ingest.dataset_from_synth(cfg)
    gpu, cpu = capi.load_product(ds), pl.oracle_engine(ds)
    for e in (gpu, cpu):
        e.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, mult_hyp, cfg.q_value, cfg.prob_cutoff)
    return ds, gpu, cpu


@pytest.mark.parametrize("mode", [capi.MH_NONE, capi.MH_BONFERRONI, capi.MH_FDR])
def test_cuda_matches_oracle_small(mode):
    cfg = synth.config_small(seed=30 + mode)
    ds, gpu, cpu = _pair(cfg, mode)
    rg, rc = gpu.level2_all_pairs(), cpu.level2_all_pairs()
    assert rg.n_multiplicons > 50
    assert capi.compare_packed(rg.bytes, rc.bytes, pl.P_RTOL) is None
    assert np.array_equal(rg.ghm_index, rc.ghm_index)
    assert rg.stats["n_dots"] == rc.stats["n_dots"]
    # profile searches with masks accumulating like higherLevel.cpp:165-171
    ms = capi.parse_packed(rc.bytes)
    order = sorted(range(len(ms)), key=lambda i: -sum(len(c.aps) for c in ms[i].clusters))[:25]
    found = 0
    for i in order:
        m = ms[i]
        prof = pl.simple_profile(ds, m.x_object_id, m.y_object_id, m)
        for e in (gpu, cpu):
            e.mask_range(m.x_object_id, m.bounds[0], m.bounds[1])
            e.mask_range(m.y_object_id, m.bounds[2], m.bounds[3])
        pg, pc = gpu.profile_search(prof), cpu.profile_search(prof)
        assert capi.compare_packed(pg.bytes, pc.bytes, pl.P_RTOL) is None, f"profile {i}"
        found += pg.n_multiplicons
    assert found > 0


def test_fdr_sort_with_many_tied_pvalues():
    """> 16 clusters with equal p-values in one GHM: exercises the restated libstdc++
    introsort (unstable on ties) on the device."""
    n = 40
    fam, gid, flg, off = [], [], [], [0]
    nf = 0
    # list 0: n blocks of 4 collinear genes separated by 60 fillers; list 1: same blocks, same spacing
    for l in range(2):
        f, g, s = [], [], []
        for b in range(n):
            for k in range(4):
                f.append(b * 4 + k)
            for k in range(60):
                f.append(100000 + nf)
                nf += 1
        fam += f
        gid += list(range(len(gid), len(gid) + len(f)))
        flg += [1] * len(f)
        off.append(len(fam))
    # densify family ids
    u = {v: i for i, v in enumerate(sorted(set(fam)))}
    ds = capi.Dataset(np.array(off, np.int64), np.array([u[v] for v in fam], np.int32), np.array(gid, np.int32),
                      np.arange(len(fam), dtype=np.int32), np.array(flg, np.uint8))
    gpu, cpu = capi.load_product(ds), pl.oracle_engine(ds)
    for e in (gpu, cpu):
        e.set_params(mult_hyp=capi.MH_FDR, prob_cutoff=0.5)
    rg, rc = gpu.level2_pairs([0], [1]), cpu.level2_pairs([0], [1])
    ms = capi.parse_packed(rc.bytes)
    assert sum(len(m.clusters) for m in ms) >= 17
    ps = [c.p for m in ms for c in m.clusters]
    assert len(set(ps)) < len(ps)  # ties present
    assert capi.compare_packed(rg.bytes, rc.bytes, pl.P_RTOL) is None


def test_edge_cases():
    # empty dataset
    ds0 = capi.Dataset(np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32),
                       np.zeros(0, np.uint8))
    e = capi.load_product(ds0)
    r = e.level2_all_pairs()
    assert r.n_multiplicons == 0 and r.bytes == b"\x00\x00\x00\x00"
    e.close()
    # ragged: empty list, 1-gene list, lists without any homolog, self comparison only
    off = np.array([0, 0, 1, 4, 9], np.int64)
    fam = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8], np.int32)
    ds1 = capi.Dataset(off, fam, np.arange(9, dtype=np.int32), np.arange(9, dtype=np.int32), np.ones(9, np.uint8))
    gpu, cpu = capi.load_product(ds1), pl.oracle_engine(ds1)
    rg, rc = gpu.level2_all_pairs(), cpu.level2_all_pairs()
    assert rg.bytes == rc.bytes == b"\x00\x00\x00\x00"
    assert rg.stats["n_ghm"] == 10
    # bad arguments are errors, not crashes
    with pytest.raises(RuntimeError):
        gpu.level2_pairs([0], [7])
    with pytest.raises(RuntimeError):
        gpu.mask_range(3, 0, 99)
    with pytest.raises(RuntimeError):
        gpu.set_params(anchor_points=2)
    gpu.close()


def test_everything_masked_profile_search_is_empty():
    tr = read_trace(pl.golden_traces()[0])
    gpu = capi.load_product(tr.dataset)
    pl.set_params_from(gpu, tr.params)
    for l in range(tr.dataset.n_lists):
        gpu.mask_range(l, 0, tr.dataset.size(l) - 1)
    prof = [ev[1] for ev in tr.events if ev[0] == "profile"][0]
    r = gpu.profile_search(prof)
    assert r.n_multiplicons == 0 and r.stats["n_ghm"] == 0
    gpu.clear_masks()
    assert gpu.profile_search(prof).stats["n_ghm"] == tr.dataset.n_lists
    gpu.close()


def test_list_longer_than_reference_limit_is_rejected():
    n = 46341
    ds = capi.Dataset(np.array([0, n], np.int64), np.zeros(n, np.int32), np.arange(n, dtype=np.int32),
                      np.arange(n, dtype=np.int32), np.ones(n, np.uint8))
    with pytest.raises(RuntimeError, match="46340"):
        capi.load_product(ds)


def test_omxn_device_known_answers():
    tr = read_trace(pl.golden_traces()[0])
    gpu, cpu = capi.load_product(tr.dataset), pl.oracle_engine(tr.dataset)
    x = np.concatenate([[1e-4, 0.3], np.linspace(1e-9, 0.6, 500)])
    n = np.concatenate([[1800, 7], np.arange(1, 501)]).astype(np.int32)
    a, b = gpu.omxn(x, n), cpu.omxn(x, n)
    assert a[0].hex() == "0x1.aba78d4067e56p-1"
    assert abs(a[1] - float.fromhex("0x1.5152be12f5a5ep-4")) <= 2 ** -56
    assert np.max(np.abs(a - b) / np.abs(b)) < 1e-15
    gpu.close()


def test_pow_matches_host_libm():
    """The device pow of the probability kernel (glibc 2.39 FMA variant restated) against the
    host libm of this box, bit for bit, over the (density, exponent) domain of
    BaseCluster::calculateProbability."""
    tr = read_trace(pl.golden_traces()[0])
    gpu, cpu = capi.load_product(tr.dataset), pl.oracle_engine(tr.dataset)
    rng = np.random.RandomState(7)
    n = 2_000_000
    dens = np.where(rng.rand(n) < 0.5, 10.0 ** (-9 + 8.7 * rng.rand(n)), rng.rand(n) * 0.999)
    base = 1.0 - dens
    c = np.maximum(1, (10.0 ** (5.3 * rng.rand(n))).astype(np.int32))
    c[::7] = 1 + rng.randint(0, 3000, size=len(c[::7]))
    a, b = gpu.pow(base, c), cpu.pow(base, c)
    big = np.abs(c * np.log(base)) >= 512     # documented: only 1 - pow() has to agree there
    assert np.array_equal(a[~big], b[~big])
    assert np.array_equal(1.0 - a[big], 1.0 - b[big])
    gpu.close()


def test_pool_exhaustion_is_retried(monkeypatch):
    """A deliberately tiny anchor-point pool must trigger the engine's rerun with a larger pool
    and still give the reference's answer."""
    tr = read_trace([p for p in pl.golden_traces() if "tiny_4g" in p][0])
    monkeypatch.setenv("IADH_POOL_FACTOR", "1")
    monkeypatch.setenv("IADH_RESIDENT", "0")
    gpu = capi.load_product(tr.dataset)
    monkeypatch.delenv("IADH_POOL_FACTOR")
    errs = pl.replay_trace(gpu, tr, rtol_p=pl.P_RTOL, check_dots=False)
    gpu.close()
    assert not errs, errs[:3]


@pytest.mark.parametrize("scale", [float(os.environ.get("IADH_TEST_C3_SCALE", "1.0"))])
def test_c3_full_size_properties(scale):
    """BASELINE config 3 (10 genomes x 30k genes, 100 lists, 5050 pairs) through the CUDA path
    only; checked through size-independent properties."""
    cfg = synth.config_c3(scale)
    ds = ingest.dataset_from_synth(cfg)
    gpu = capi.load_product(ds)
    gpu.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    r = gpu.level2_all_pairs()
    n = ds.n_lists
    assert r.stats["n_ghm"] == n * (n + 1) // 2
    # idempotence
    r2 = gpu.level2_all_pairs()
    assert r2.bytes == r.bytes
    # sharding: the union of the ranks' streams is the whole stream, GHM by GHM
    xs, ys = capi.plan_pairs(np.diff(ds.list_off))
    key = {(int(x), int(y)): i for i, (x, y) in enumerate(zip(xs, ys))}
    ms = capi.parse_packed(r.bytes)
    per_ghm = {}
    for m, g in zip(ms, r.ghm_index):
        per_ghm.setdefault(int(g), []).append(m)
    shard_total = 0
    for rank in range(2):
        rs = gpu.level2_all_pairs(rank, 2)
        xr, yr = capi.plan_pairs(np.diff(ds.list_off), rank, 2)
        msr = capi.parse_packed(rs.bytes)
        shard_total += len(msr)
        for m, g in zip(msr, rs.ghm_index):
            full = per_ghm[key[(int(xr[g]), int(yr[g]))]]
            assert any(np.array_equal(m.clusters[0].aps, f.clusters[0].aps) for f in full)
    assert shard_total == len(ms)
    # every anchor point is a homologous pair of different genes; clusters obey the thresholds
    nap = 0
    for m in ms:
        xo, yo = ds.list_off[m.x_object_id], ds.list_off[m.y_object_id]
        assert m.level == 2 and sum(len(c.aps) for c in m.clusters) >= cfg.anchor_points
        for c in m.clusters:
            a = c.aps
            nap += len(a)
            assert np.all(np.diff(a[:, 2]) >= 0)                       # multiset order: x ascending
            assert np.all(ds.fam[xo + a[:, 2]] == ds.fam[yo + a[:, 3]])
            assert np.all(ds.gid[xo + a[:, 2]] != ds.gid[yo + a[:, 3]])
            assert np.all(ds.num_id[xo + a[:, 2]] == a[:, 0]) and np.all(ds.num_id[yo + a[:, 3]] == a[:, 1])
            same = (ds.flags[xo + a[:, 2]] & 1) == (ds.flags[yo + a[:, 3]] & 1)
            assert c.bounds == (a[:, 2].min(), a[:, 2].max(), a[:, 3].min(), a[:, 3].max())
            assert 0.0 <= c.p <= cfg.prob_cutoff
            x, y = a[:, 2].astype(float), a[:, 3].astype(float)
            r2v = np.corrcoef(x, y)[0, 1] ** 2
            assert r2v >= cfg.q_value - 1e-9
            assert (np.corrcoef(x, y)[0, 1] > 0) == bool(c.orient)
    assert nap == r.stats["n_anchorpoints"] and len(ms) == r.stats["n_multiplicons"]
    gpu.close()


def _dup_id_dataset(seed=5):
    """Small lists where gene IDs repeat inside a list and across lists (the reference compares ID
    strings, GHM.cpp:68 / Gene.cpp:20): exercises the per-lane slow path of the level-2 build."""
    rng = np.random.RandomState(seed)
    sizes = [180, 150, 1, 0, 120]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(off[-1])
    fam = rng.randint(0, 25, n).astype(np.int32)          # high-copy families: dense GHMs
    gid = np.arange(n, dtype=np.int32)
    dup = rng.choice(n, 60, replace=False)
    gid[dup] = gid[rng.choice(n, 60)]                      # same ID in several places (same or other family)
    fam[dup[:30]] = fam[gid[dup[:30]]]                     # ... and half of them homologous to their namesake
    flags = rng.randint(0, 2, n).astype(np.uint8)
    return capi.Dataset(off, fam, gid, np.arange(n, dtype=np.int32) + 1000, flags)


@pytest.mark.parametrize("host_tables", [False, True])
@pytest.mark.parametrize("no_table", [False, True])
def test_build_with_duplicate_gene_ids(no_table, host_tables, monkeypatch):
    ds = _dup_id_dataset()
    if no_table:
        monkeypatch.setenv("IADH_NO_FAM_TABLE", "1")
    if host_tables:   # iadh_create builds the join tables on the device by default; this is the host construction
        monkeypatch.setenv("IADH_HOST_TABLES", "1")
    gpu, cpu = capi.load_product(ds), pl.oracle_engine(ds)
    xs, ys = capi.plan_pairs(np.diff(ds.list_off))
    dg, dc = gpu.build_dots(xs, ys), cpu.build_dots(xs, ys)
    assert sum(len(a) + len(b) for a, b in dc) > 2000
    for (a0, a1), (b0, b1) in zip(dg, dc):
        assert np.array_equal(a0, b0) and np.array_equal(a1, b1)
    st = gpu.build_only(xs, ys)
    assert st["n_dots"] == sum(len(a) + len(b) for a, b in dc) and st["build_ms"] > 0
    for e in (gpu, cpu):
        e.set_params(mult_hyp=capi.MH_NONE, prob_cutoff=0.9)
    assert capi.compare_packed(gpu.level2_all_pairs().bytes, cpu.level2_all_pairs().bytes, pl.P_RTOL) is None
    gpu.close()


@pytest.mark.parametrize("sort_max", [None, "64"])
def test_family_expansion_build_is_invisible(sort_max, monkeypatch):
    """K1's family-expansion path (count / append / block sort; taken by level-2 batches of >= 1,024 GHMs) forced on
    small inputs: same dots in the same order as the oracle, same packed bytes -- family and pairs mode, repeated gene
    IDs, pair lists in any orientation, a pair named twice (tile kernels), and blocks beyond the sort's capacity
    (sort_max = 64: the call falls back to the tile kernels after the count pass)."""
    monkeypatch.setenv("IADH_EXPAND_MIN", "1")
    monkeypatch.setenv("IADH_HL_EXPAND_MIN", "1")
    if sort_max:
        monkeypatch.setenv("IADH_EXPAND_SORT_MAX", sort_max)
    sets = [_dup_id_dataset(), ingest.dataset_from_synth(synth.config_small(seed=41)), ingest.dataset_from_synth(synth.config_c5(0.03))]
    cfgp = synth.config_small(seed=23)
    cfgp.table_type = "pairs"
    sets.append(ingest.dataset_from_synth(cfgp))
    for k, ds in enumerate(sets):
        gpu, cpu = capi.load_product(ds), pl.oracle_engine(ds)
        xs, ys = capi.plan_pairs(np.diff(ds.list_off))
        lists = [(xs, ys), (list(ys[:40]) + [1, 0], list(xs[:40]) + [1, 0]), ([0, 2, 0], [1, 2, 1])]
        for xa, ya in lists:
            dg, dc = gpu.build_dots(xa, ya), cpu.build_dots(xa, ya)
            assert len(dg) == len(dc) and sum(len(a) + len(b) for a, b in dc) > 100
            for (a0, a1), (b0, b1) in zip(dg, dc):
                assert np.array_equal(a0, b0) and np.array_equal(a1, b1), k
        if k != 2:   # (the dense shape is a build-only workload)
            for e in (gpu, cpu):
                e.set_params(mult_hyp=capi.MH_BONFERRONI)
            a, b = gpu.level2_all_pairs(), cpu.level2_all_pairs()
            assert capi.compare_packed(a.bytes, b.bytes, pl.P_RTOL) is None
            a, b = gpu.level2_all_pairs(1, 3), cpu.level2_all_pairs(1, 3)     # one rank's share of the pairs
            assert capi.compare_packed(a.bytes, b.bytes, pl.P_RTOL) is None
            # profile GHMs (family mode: expansion over the column groups; pairs mode: tile kernel), masks accumulating
            ms = capi.parse_packed(cpu.level2_all_pairs().bytes)
            order = sorted(range(len(ms)), key=lambda i: -sum(len(c.aps) for c in ms[i].clusters))[:6]
            for i in order:
                m = ms[i]
                prof = pl.simple_profile(ds, m.x_object_id, m.y_object_id, m)
                for e in (gpu, cpu):
                    e.mask_range(m.x_object_id, m.bounds[0], m.bounds[1])
                    e.mask_range(m.y_object_id, m.bounds[2], m.bounds[3])
                ly = list(range(ds.n_lists))
                da, db = gpu.build_dots([0] * ds.n_lists, ly, profile=prof), cpu.build_dots([0] * ds.n_lists, ly, profile=prof)
                assert all(np.array_equal(u[0], v[0]) and np.array_equal(u[1], v[1]) for u, v in zip(da, db)), (k, i)
                assert capi.compare_packed(gpu.profile_search(prof).bytes, cpu.profile_search(prof).bytes, pl.P_RTOL) is None
        gpu.close()


def test_result_exchange_with_a_single_rank_is_the_identity():
    """iadh_comm_init / iadh_allgather_results with a communicator of one rank (what a 1-GPU box can run of the
    multi-GPU exchange; 2-8 ranks: tools/dist_check.py under torchrun): the gathered streams equal the rank's own, for
    a single result and for the results of a batch (slices of one pinned block, counts written on the device)."""
    cfg = synth.config_small(seed=30)
    ds = ingest.dataset_from_synth(cfg)
    eng = capi.load_product(ds)
    eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    eng.comm_init(eng.comm_unique_id(), 0, 1)
    r = eng.level2_all_pairs()
    g = eng.allgather([r])[0]
    assert g.n_multiplicons == r.n_multiplicons > 50 and g.bytes == r.bytes
    ms = capi.parse_packed(g.bytes)
    order = sorted(range(len(ms)), key=lambda i: -sum(len(c.aps) for c in ms[i].clusters))[:6]
    profs = [pl.simple_profile(ds, ms[i].x_object_id, ms[i].y_object_id, ms[i]) for i in order]
    masks = [[(ms[i].x_object_id, ms[i].bounds[0], ms[i].bounds[1]), (ms[i].y_object_id, ms[i].bounds[2], ms[i].bounds[3])] for i in order]
    res = eng.profile_search_batch(profs, masks)
    got = eng.allgather(res)
    assert sum(x.n_multiplicons for x in res) > 0
    assert all(a.bytes == b.bytes for a, b in zip(got, res))
    eng.close()


def test_device_pairs_tandem_remap_matches_reference_and_oracle():
    """iadh_remap_tandems_pairs (GeneList::remapTandemsPairs with indirect pairs, on the device) against the reference's
    own flags in the table_type=pairs traces and against the oracle on a larger pairs-mode set."""
    for path in pl.golden_traces("pairs"):
        tr = read_trace(path)
        off = np.concatenate([[0], np.cumsum([len(r) for r in tr.raw_lists])]).astype(np.int64)
        gid = np.concatenate([r["gid"] for r in tr.raw_lists]).astype(np.int32)
        strand = np.concatenate([r["strand"] for r in tr.raw_lists]).astype(np.uint8)
        ref = np.concatenate([r["remapped"] for r in tr.raw_lists]).astype(bool)
        got = capi.remap_tandems_pairs(off, gid, strand, tr.dataset.pair_ptr, tr.dataset.pair_ids, tr.params["tandem_gap"])
        assert ref.sum() > 0 and np.array_equal(got, ref), path
    cfg = synth.config_small(seed=23)
    cfg.table_type = "pairs"
    data = synth.generate(cfg)
    lists = [lst for g in cfg.genomes for lst in data[g.name]]
    gid_map = {}
    ps = ingest.pair_sets(lists, synth.pair_table(cfg, data), gid_map)
    pp, pi = ingest.pairs_csr(ps, len(gid_map))
    off = np.concatenate([[0], np.cumsum([len(l) for l in lists])]).astype(np.int64)
    gid = np.array([gid_map[g[0]] for l in lists for g in l], np.int32)
    strand = np.array([g[1] for l in lists for g in l], np.uint8)
    tg = cfg.tandem_gap if cfg.tandem_gap is not None else cfg.gap_size // 2
    a = capi.remap_tandems_pairs(off, gid, strand, pp, pi, tg)
    b = capi.remap_tandems_pairs(off, gid, strand, pp, pi, tg, lib_path=capi.ORACLE_LIB, prefix="oracle_")
    assert b.sum() > 50 and np.array_equal(a, b)
    # and the ingest mirror with the device path gives the dataset of the pure-Python path
    d0, d1 = ingest.dataset_from_synth(cfg, data=data), ingest.dataset_from_synth(cfg, data=data, device=0)
    for k in ("list_off", "gid", "num_id", "flags", "pair_ptr", "pair_ids"):
        assert np.array_equal(getattr(d0, k), getattr(d1, k)), k


def test_device_string_interning_matches_oracle():
    """iadh_intern_strings (the hashing half of the family join, on the device) against the oracle's unordered_map:
    same dense ids in first-occurrence order -- gene and family names of a synthetic set, heavy duplication, empty and
    long strings, strings that differ only in their last byte, and 300k strings over 50k distinct values."""
    rng = np.random.RandomState(5)
    cfg = synth.config_small(seed=9)
    data = synth.generate(cfg)
    genes = [g for gn in cfg.genomes for lst in data[gn.name] for g in lst]
    cases = [[g[0] for g in genes] + [g[0] for g in genes[::3]],
             ["F%d" % g[2] for g in genes],
             ["", "a", "", "a" * 1000, "a" * 999 + "b", "a" * 1000, "b", "a"],
             ["fam%d" % v for v in rng.randint(0, 50000, 300000)],
             ["same"] * 5000]
    for strings in cases:
        a, na = capi.intern_strings(strings)
        b, nb = capi.intern_strings(strings, lib_path=capi.ORACLE_LIB, prefix="oracle_")
        assert na == nb == len(set(strings))
        assert np.array_equal(a, b)
    a, na = capi.intern_strings([])
    assert len(a) == 0 and na == 0


@pytest.mark.parametrize("host_tables", [False, True])
def test_join_tables_with_giant_families(host_tables, monkeypatch):
    """iadh_create's device-built join tables on family sizes that take every path of the bucket sort: 1 ... 32 members
    (one register per lane), hundreds (8 / 32 registers), 3,000 (shared memory) and 10,000 (global-memory network with
    virtual padding) -- dots of a few list pairs against the oracle, with the host construction as a second witness."""
    if host_tables:
        monkeypatch.setenv("IADH_HOST_TABLES", "1")
    rng = np.random.RandomState(17)
    sizes = [3000, 10000] + [700, 300, 150, 40] * 3 + list(rng.randint(1, 33, 400))
    fam = np.repeat(np.arange(len(sizes), dtype=np.int32), sizes)
    rng.shuffle(fam)
    n = len(fam)
    n_lists = 12
    off = np.linspace(0, n, n_lists + 1).astype(np.int64)
    gid = np.arange(n, dtype=np.int32)
    flags = rng.randint(0, 2, n).astype(np.uint8)
    ds = capi.Dataset(off, fam, gid, np.arange(n, dtype=np.int32) + 7, flags)
    gpu, cpu = capi.load_product(ds), pl.oracle_engine(ds)
    xs, ys = [0, 3, 5, 11, 2], [0, 4, 5, 1, 9]
    dg, dc = gpu.build_dots(xs, ys), cpu.build_dots(xs, ys)
    assert sum(len(a) + len(b) for a, b in dc) > 10 ** 6
    for (a0, a1), (b0, b1) in zip(dg, dc):
        assert np.array_equal(a0, b0) and np.array_equal(a1, b1)
    gpu.close()


def test_dense_build_matches_oracle():
    """BASELINE config 5 shape at 3 % scale (tandem-rich, >= 50 members per family): dots of every
    list pair, both orientations, in the reference's (x, y) order."""
    cfg = synth.config_c5(0.03)
    ds = ingest.dataset_from_synth(cfg)
    gpu, cpu = capi.load_product(ds), pl.oracle_engine(ds)
    xs, ys = capi.plan_pairs(np.diff(ds.list_off))
    dg, dc = gpu.build_dots(xs, ys), cpu.build_dots(xs, ys)
    assert sum(len(a) + len(b) for a, b in dc) > 100000
    for (a0, a1), (b0, b1) in zip(dg, dc):
        assert np.array_equal(a0, b0) and np.array_equal(a1, b1)
    gpu.close()


@pytest.mark.parametrize("split_min", ["3", "1000000"])
def test_orientation_split_is_invisible(split_min, monkeypatch):
    """The two orientations of a GHM may be clustered by two CTAs (IADH_SPLIT_MIN): same bytes."""
    tr = read_trace([p for p in pl.golden_traces() if "tiny_4g" in p][0])
    monkeypatch.setenv("IADH_RESIDENT", "0")     # (the split belongs to the global-memory kernels)
    monkeypatch.setenv("IADH_SPLIT_MIN", split_min)
    gpu = capi.load_product(tr.dataset)
    errs = pl.replay_trace(gpu, tr, rtol_p=pl.P_RTOL)
    gpu.close()
    assert not errs, errs[:3]
    cfg = synth.config_small(seed=41)
    ds, gpu, cpu = _pair(cfg, capi.MH_BONFERRONI)
    assert gpu.level2_all_pairs().bytes == cpu.level2_all_pairs().bytes
    gpu.close()


def test_event_hooks_time_the_engine_stream():
    tr = read_trace(pl.golden_traces()[0])
    gpu = capi.load_product(tr.dataset)
    gpu.event_record(0)
    r = gpu.level2_all_pairs()
    gpu.event_record(1)
    ms = gpu.event_elapsed_ms(0, 1)
    assert ms >= r.stats["cluster_ms"] > 0
    with pytest.raises(RuntimeError):
        gpu.event_record(9)
    gpu.close()


def test_device_tandem_remap_matches_reference_and_oracle():
    """iadh_remap_tandems (GeneList.cpp:143-200 on the device) against the reference's own flags in
    the golden traces, and against the oracle on tandem-rich lists (config 5 shape, tandem_gap 10)
    plus ragged / empty / single-gene lists and repeated gene IDs."""
    for path in pl.golden_traces("family"):
        tr = read_trace(path)
        off = np.concatenate([[0], np.cumsum([len(r) for r in tr.raw_lists])]).astype(np.int64)
        cat = lambda k: np.concatenate([r[k] for r in tr.raw_lists])
        got = capi.remap_tandems(off, cat("fam"), cat("gid"), cat("strand"), tr.params["tandem_gap"])
        assert np.array_equal(got, cat("remapped").astype(bool)), path
    cfg = synth.config_c5(0.1)
    data = synth.generate(cfg)
    lists = [lst for g in cfg.genomes for lst in data[g.name]] + [[], [("solo", True, 0)]]
    off = np.concatenate([[0], np.cumsum([len(l) for l in lists])]).astype(np.int64)
    fam = np.array([g[2] for l in lists for g in l], np.int32)
    strand = np.array([g[1] for l in lists for g in l], np.uint8)
    gid = np.arange(len(fam), dtype=np.int32)
    gid[5::97] = gid[4::97][:len(gid[5::97])]          # neighbours sharing one gene ID never pair
    for gap in (0, 1, 10, 15):
        got = capi.remap_tandems(off, fam, gid, strand, gap)
        want = capi.remap_tandems(off, fam, gid, strand, gap, lib_path=capi.ORACLE_LIB, prefix="oracle_")
        assert np.array_equal(got, want), gap
    assert want.sum() > 1000


@pytest.mark.parametrize("cfg_code", [None, "3232:32", "3224:24", "32:16"])
def test_pass_per_launch_mode_is_invisible(cfg_code, monkeypatch):
    """Phase 1 as one pass per launch (the path of large throughput-bound batches) forced on small
    inputs, level 2 and profile searches: same bytes as the fused kernel / the reference -- with the launch
    configuration the engine picks for the batch size and with each of the three it chooses from."""
    if cfg_code:
        monkeypatch.setenv("IADH_PASS_CFG", cfg_code)
    monkeypatch.setenv("IADH_PASS_MIN", "1")
    monkeypatch.setenv("IADH_PASS_RATIO", "0")
    monkeypatch.setenv("IADH_RESIDENT", "0")     # (the path of GHMs too large for the resident kernel)
    tr = read_trace([p for p in pl.golden_traces() if "tiny_4g" in p][0])
    gpu = capi.load_product(tr.dataset)
    errs = pl.replay_trace(gpu, tr, rtol_p=pl.P_RTOL)
    gpu.close()
    assert not errs, errs[:3]
    for mode in (capi.MH_BONFERRONI, capi.MH_FDR):
        cfg = synth.config_small(seed=50 + mode)
        ds, gpu, cpu = _pair(cfg, mode)
        rg = gpu.level2_all_pairs()
        assert rg.stats["kernel_launches"] > 30          # 10 gaps x 3 passes + phase 2 + build + pack
        assert rg.bytes == cpu.level2_all_pairs().bytes
        gpu.close()


def test_pass_per_launch_mode_retries_pool_exhaustion(monkeypatch):
    """The rerun with a larger anchor-point pool also works when the state travels between launches."""
    monkeypatch.setenv("IADH_PASS_MIN", "1")
    monkeypatch.setenv("IADH_PASS_RATIO", "0")
    monkeypatch.setenv("IADH_RESIDENT", "0")
    monkeypatch.setenv("IADH_POOL_FACTOR", "1")
    tr = read_trace([p for p in pl.golden_traces() if "tiny_4g" in p][0])
    gpu = capi.load_product(tr.dataset)
    errs = pl.replay_trace(gpu, tr, rtol_p=pl.P_RTOL, check_dots=False)
    gpu.close()
    assert not errs, errs[:3]


def test_level2_batches_are_bisected_to_fit_memory(monkeypatch):
    """A level-2 call whose clustering workspace exceeds the budget is bisected into consecutive
    ranges of the pair sequence (IADH_WS_BUDGET_MB forces it here): same stream, same GHM indices."""
    cfg = synth.config_small(seed=61)
    ds, gpu, cpu = _pair(cfg, capi.MH_BONFERRONI)
    want = cpu.level2_all_pairs()
    whole = gpu.level2_all_pairs()
    gpu.close()
    monkeypatch.setenv("IADH_WS_BUDGET_MB", "1")
    gpu2 = capi.load_product(ds)
    gpu2.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    split = gpu2.level2_all_pairs()
    assert split.stats["kernel_launches"] > 3 * whole.stats["kernel_launches"]     # really split
    assert split.bytes == whole.bytes == want.bytes
    assert np.array_equal(split.ghm_index, want.ghm_index)
    for k in ("n_ghm", "n_dots", "n_multiplicons", "n_anchorpoints"):
        assert split.stats[k] == whole.stats[k]
    gpu2.close()


def test_ingest_with_device_tandem_remap_equals_the_host_mirror():
    """The dataset bench.py builds (tandem remapping on the device) is the one the CPU tests build."""
    for cfg in (synth.config_small(seed=71), synth.config_c5(0.02)):
        a = ingest.dataset_from_synth(cfg)
        b = ingest.dataset_from_synth(cfg, device=0)
        for k in ("list_off", "fam", "gid", "num_id", "flags"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), (cfg.name, k)


def test_pairs_mode_matches_oracle_and_refuses_masked_level2():
    """table_type=pairs (iadh_create_pairs) beyond the golden traces: a larger set with the (gene, list)
    table and with the binary-search fallback, all three filter modes, profile searches with masks; and
    the one documented difference: a level-2 call with masked elements is refused."""
    cfg = synth.config_small(seed=23)
    cfg.table_type = "pairs"
    ds = ingest.dataset_from_synth(cfg)
    assert ds.pairs_mode
    cpu = pl.oracle_engine(ds)
    for no_table in (False, True):
        if no_table:
            os.environ["IADH_NO_FAM_TABLE"] = "1"
        try:
            gpu = capi.load_product(ds)
        finally:
            os.environ.pop("IADH_NO_FAM_TABLE", None)
        for mode in (capi.MH_BONFERRONI, capi.MH_FDR, capi.MH_NONE):
            for e in (gpu, cpu):
                e.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, mode, cfg.q_value, cfg.prob_cutoff)
                e.clear_masks()
            xs, ys = [0, 1, 2, 3, 5, 7, 4], [0, 0, 1, 3, 2, 7, 9]
            a, b = gpu.level2_pairs(xs, ys), cpu.level2_pairs(xs, ys)
            assert a.n_multiplicons > 0 and not capi.compare_packed(a.bytes, b.bytes, pl.P_RTOL)
            da, db = gpu.build_dots(xs, ys), cpu.build_dots(xs, ys)
            assert all(np.array_equal(u[0], v[0]) and np.array_equal(u[1], v[1]) for u, v in zip(da, db))
            if no_table and mode != capi.MH_BONFERRONI:
                continue
            found = 0
            ms = capi.parse_packed(b.bytes)
            order = sorted(range(len(ms)), key=lambda i: -sum(len(c.aps) for c in ms[i].clusters))[:8]
            for i in order:
                m = ms[i]
                prof = pl.simple_profile(ds, m.x_object_id, m.y_object_id, m)
                for e in (gpu, cpu):
                    e.mask_range(m.x_object_id, m.bounds[0], m.bounds[1])
                    e.mask_range(m.y_object_id, m.bounds[2], m.bounds[3])
                ra, rb = gpu.profile_search(prof), cpu.profile_search(prof)
                assert not capi.compare_packed(ra.bytes, rb.bytes, pl.P_RTOL), i
                da = gpu.build_dots([0] * ds.n_lists, list(range(ds.n_lists)), profile=prof)
                db = cpu.build_dots([0] * ds.n_lists, list(range(ds.n_lists)), profile=prof)
                assert all(np.array_equal(u[0], v[0]) and np.array_equal(u[1], v[1]) for u, v in zip(da, db))
                found += ra.n_multiplicons
            assert found > 0
        gpu.mask_range(1, 0, 5)
        with pytest.raises(RuntimeError, match="masked"):
            gpu.level2_pairs([0], [1])
        gpu.close()


@pytest.mark.parametrize("knobs", [{}, {"IADH_RES_WIDE": "1:64"}, {"IADH_RES_WIDE": "1:256"}, {"IADH_RES_SQUEEZE": "1"},
                                   {"IADH_RES_SQUEEZE": "1", "IADH_POOL_FACTOR": "1"}, {"IADH_RESIDENT": "0"}, {"IADH_JOIN2": "0"}],
                         ids=["resident", "wide64", "wide256", "squeezed", "squeezed_pool1", "global_only", "global_rowjoin"])
def test_resident_kernel_variants_are_invisible(knobs, monkeypatch):
    """The resident kernel (a GHM's clustering state in the shared memory of one CTA) on 32-, 64- and
    256-thread CTAs, with budgets so small that most GHMs outgrow them and are re-run on the global path
    (IADH_RES_SQUEEZE; with IADH_POOL_FACTOR=1 that re-run itself retries with a larger pool), and switched
    off: the reference's bytes every time, level 2 and profile searches, all three filter modes."""
    monkeypatch.setenv("IADH_RESIDENT", "1")
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    if "IADH_JOIN2" in knobs:          # the row-by-row join of the pass-per-launch path (the two-phase join is the default)
        monkeypatch.setenv("IADH_RESIDENT", "0")
        monkeypatch.setenv("IADH_PASS_MIN", "1")
        monkeypatch.setenv("IADH_PASS_RATIO", "0")
    for name in ("tiny_4g", "tiny_fdr", "tiny_none", "tiny_pairs"):
        tr = read_trace([p for p in pl.golden_traces() if name in p][0])
        gpu = capi.load_product(tr.dataset)
        errs = pl.replay_trace(gpu, tr, rtol_p=pl.P_RTOL, check_dots=False)
        gpu.close()
        assert not errs, (name, errs[:3])
    cfg = synth.config_small(seed=77)
    ds, gpu, cpu = _pair(cfg, capi.MH_BONFERRONI)
    assert gpu.level2_all_pairs().bytes == cpu.level2_all_pairs().bytes
    gpu.close()


@pytest.mark.parametrize("table_type", ["family", "pairs"])
def test_profile_search_batch_equals_sequential_calls(table_type):
    """iadh_profile_search_batch (K profiles, the masks of set k applied before profile k, every GHM of the
    batch in the same launches) against the same searches one call at a time on the device AND against the
    oracle's sequential calls; then the list-sharded and the profile-sharded forms over 3 ranks; the masks
    left in the context afterwards are those of the sequential calls."""
    cfg = synth.config_small(seed=91)
    cfg.table_type = table_type
    ds, gpu, cpu = _pair(cfg, capi.MH_BONFERRONI)
    ms = capi.parse_packed(cpu.level2_all_pairs().bytes)
    order = sorted(range(len(ms)), key=lambda i: -sum(len(c.aps) for c in ms[i].clusters))[:12]
    profs = [pl.simple_profile(ds, ms[i].x_object_id, ms[i].y_object_id, ms[i]) for i in order]
    masks = [[(ms[i].x_object_id, ms[i].bounds[0], ms[i].bounds[1]), (ms[i].y_object_id, ms[i].bounds[2], ms[i].bounds[3])]
             for i in order]
    masks[3] = []                       # a profile without new masks
    masks[5] = masks[5] + masks[5]      # overlapping ranges
    want = []
    for p, mk in zip(profs, masks):
        for (l, b, e) in mk:
            cpu.mask_range(l, b, e)
        want.append(cpu.profile_search(p))
    assert sum(w.n_multiplicons for w in want) > 0
    # batch on the device, all lists, one rank
    got = gpu.profile_search_batch(profs, masks)
    for k, (g, w) in enumerate(zip(got, want)):
        assert capi.compare_packed(g.bytes, w.bytes, pl.P_RTOL) is None, k
        assert np.array_equal(g.ghm_index, w.ghm_index) and g.stats["n_ghm"] == w.stats["n_ghm"], k
    assert got[0].stats["kernel_launches"] > 0
    # the context now holds all masks: one more sequential search sees them
    extra = pl.simple_profile(ds, ms[order[0]].x_object_id, ms[order[0]].y_object_id, ms[order[0]])
    assert capi.compare_packed(gpu.profile_search(extra).bytes, cpu.profile_search(extra).bytes, pl.P_RTOL) is None
    # sequential device calls from a clean context give the same bytes as the batch
    gpu.clear_masks()
    for k, (p, mk) in enumerate(zip(profs, masks)):
        for (l, b, e) in mk:
            gpu.mask_range(l, b, e)
        assert gpu.profile_search(p).bytes == got[k].bytes, k
    # sharded over 3 ranks: by lists (each rank's part of every profile) and by profiles (whole profiles round-robin)
    for shard_profiles in (False, True):
        for rank in range(3):
            gpu.clear_masks()
            cpu.clear_masks()
            a = gpu.profile_search_batch(profs, masks, rank, 3, shard_profiles)
            b = cpu.profile_search_batch(profs, masks, rank, 3, shard_profiles)
            for k in range(len(profs)):
                assert capi.compare_packed(a[k].bytes, b[k].bytes, pl.P_RTOL) is None, (shard_profiles, rank, k)
                if shard_profiles:
                    assert (a[k].stats["n_ghm"] > 0) == (k % 3 == rank)
                    if k % 3 == rank:
                        assert a[k].bytes == got[k].bytes
    gpu.close()
