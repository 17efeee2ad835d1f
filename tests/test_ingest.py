"""Host ingest (iadhore_b200/ingest.py) against the reference's own tandem remapping as
recorded in the golden traces (GeneList::remapTandemsFamily)."""
import numpy as np
import pytest

import parity_lib as pl
from iadhore_b200 import capi, ingest, synth
from trace_io import read_trace


@pytest.mark.parametrize("path", pl.golden_traces(), ids=lambda p: p.split("/")[-1].split(".")[0])
def test_tandem_remap_matches_reference(path):
    tr = read_trace(path)
    n_remapped = 0
    ds = tr.dataset
    ps = None
    if ds.pairs_mode:   # GeneList::remapTandemsPairs with indirect pairs (GeneList.cpp:83-141, Gene.cpp:33-49)
        ps = {g: set(ds.pair_ids[ds.pair_ptr[g]:ds.pair_ptr[g + 1]].tolist()) for g in range(len(ds.pair_ptr) - 1)
              if ds.pair_ptr[g + 1] > ds.pair_ptr[g]}
    for raw in tr.raw_lists:
        strand = [bool(s) for s in raw["strand"]]
        if ps is not None:
            r = ingest.remap_tandems_pairs(raw["gid"].tolist(), strand, tr.params["tandem_gap"], ps)
        else:
            r = ingest.remap_tandems_family(raw["fam"].tolist(), raw["gid"].tolist(), strand, tr.params["tandem_gap"])
        assert np.array_equal(r, raw["remapped"].astype(bool))
        n_remapped += int(r.sum())
    assert n_remapped > 0


def test_dataset_from_synth_matches_reference_numbering():
    tr = read_trace([p for p in pl.golden_traces() if "tiny_bf" in p][0])
    ds = ingest.dataset_from_synth(synth.config_tiny(seed=11))
    assert np.array_equal(ds.list_off, tr.dataset.list_off)
    assert np.array_equal(ds.num_id, tr.dataset.num_id)
    assert np.array_equal(ds.flags, tr.dataset.flags)
    # family / gene ids are interned differently; the partition they induce must be the same
    for a, b in ((ds.fam, tr.dataset.fam), (ds.gid, tr.dataset.gid)):
        m = {}
        for u, v in zip(a.tolist(), b.tolist()):
            assert m.setdefault(u, v) == v


def test_generator_is_deterministic():
    a = synth.generate(synth.config_tiny(seed=5))
    b = synth.generate(synth.config_tiny(seed=5))
    assert a == b
    r = synth.SplitMix64(1234567)
    assert [r.next() for _ in range(3)] == [6457827717110365317, 3203168211198807973, 9817491932198370423]


def _raw(tr):
    off = np.concatenate([[0], np.cumsum([len(r) for r in tr.raw_lists])]).astype(np.int64)
    cat = lambda k: np.concatenate([r[k] for r in tr.raw_lists])
    return off, cat("fam"), cat("gid"), cat("strand"), cat("remapped").astype(bool)


def test_pairs_ingest_matches_reference():
    """table_type=pairs: pair sets (GenePairs + DataSet::getGenePairs), tandem remapping and numbering of
    the ingest mirror against what the reference built for the same files."""
    tr = read_trace([p for p in pl.golden_traces("pairs") if p.endswith("tiny_pairs.trace.gz")][0])
    ds = ingest.dataset_from_synth(synth.config_tiny_pairs(seed=15))
    for k in ("list_off", "gid", "num_id", "flags", "pair_ptr", "pair_ids"):
        assert np.array_equal(getattr(ds, k), getattr(tr.dataset, k)), k
    # the relation is not family-shaped: some gene has two partners that are not partners of each other
    ps = {g: set(ds.pair_ids[ds.pair_ptr[g]:ds.pair_ptr[g + 1]].tolist()) for g in range(len(ds.pair_ptr) - 1)}
    assert any(b not in ps[a] for g in ps for a in ps[g] for b in ps[g] if a != b)
    assert all(g in ps[p] for g in ps for p in ps[g])   # symmetric


@pytest.mark.parametrize("path", pl.golden_traces("family"), ids=lambda p: p.split("/")[-1].split(".")[0])
def test_oracle_tandem_remap_matches_reference(path):
    """oracle_remap_tandems (the checker of the device kernel) against the reference's own flags."""
    tr = read_trace(path)
    off, fam, gid, strand, ref = _raw(tr)
    got = capi.remap_tandems(off, fam, gid, strand, tr.params["tandem_gap"], lib_path=capi.ORACLE_LIB, prefix="oracle_")
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("path", pl.golden_traces("pairs"), ids=lambda p: p.split("/")[-1].split(".")[0])
def test_oracle_pairs_tandem_remap_matches_reference(path):
    """oracle_remap_tandems_pairs (the checker of the device kernel) against the flags the reference's
    GeneList::remapTandemsPairs left in the table_type=pairs traces."""
    tr = read_trace(path)
    off, fam, gid, strand, ref = _raw(tr)
    ds = tr.dataset
    got = capi.remap_tandems_pairs(off, gid, strand, ds.pair_ptr, ds.pair_ids, tr.params["tandem_gap"], lib_path=capi.ORACLE_LIB,
                                   prefix="oracle_")
    assert ref.sum() > 0 and np.array_equal(got, ref)


def test_oracle_tandem_remap_matches_python_mirror_on_tandem_rich_lists():
    cfg = synth.config_c5(0.02)
    data = synth.generate(cfg)
    lists = [lst for g in cfg.genomes for lst in data[g.name]]
    off = np.concatenate([[0], np.cumsum([len(l) for l in lists])]).astype(np.int64)
    fam = np.array([g[2] for l in lists for g in l], np.int32)
    strand = np.array([g[1] for l in lists for g in l], np.uint8)
    gid = np.arange(len(fam), dtype=np.int32)
    got = capi.remap_tandems(off, fam, gid, strand, cfg.tandem_gap, lib_path=capi.ORACLE_LIB, prefix="oracle_")
    want = np.concatenate([ingest.remap_tandems_family(fam[a:b].tolist(), gid[a:b].tolist(), strand[a:b].astype(bool).tolist(),
                                                        cfg.tandem_gap) for a, b in zip(off[:-1], off[1:])])
    assert got.sum() > 100 and np.array_equal(got, want)


def test_reference_profile_dump_loads_to_the_trace_profiles():
    """profiles.load_reference_profiles (what bench.py replays on the device: the reference's own aligned
    profiles, dumped by ref_harness --dump-profiles) against the profile / mask events of the full trace of
    the same reference run."""
    import os
    import parity_lib as pl
    from iadhore_b200 import profiles
    from trace_io import read_trace
    tr = read_trace(os.path.join(pl.GOLDEN, "tiny_4g.trace.gz"))
    got = profiles.load_reference_profiles(tr.dataset, os.path.join(pl.GOLDEN, "tiny_4g.profiles.bin.gz"))
    want, masks = [], []
    for ev in tr.events:
        if ev[0] == "mask":
            masks.append((ev[1], ev[2], ev[3]))
        elif ev[0] == "profile":
            want.append((masks, ev[1]))
            masks = []
    assert len(got) == len(want) >= 20
    for (gm, gp), (wm, wp) in zip(got, want):
        assert [tuple(m) for m in gm] == wm
        assert (gp.S, gp.L, gp.x_object_id) == (wp.S, wp.L, wp.x_object_id)
        gap = (np.asarray(wp.flags) & 2) != 0
        assert np.array_equal(np.asarray(gp.flags), np.asarray(wp.flags))
        for k in ("fam", "gid", "num_id"):
            assert np.array_equal(np.asarray(getattr(gp, k))[~gap], np.asarray(getattr(wp, k))[~gap]), k
