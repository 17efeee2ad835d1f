"""Writes the input of tests/emu/emu_join2 from a reference trace: the dot sets of its level-2 and
profile GHMs (the reference's own buildMatrix output) and the gap-size schedule."""
import struct
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 3)[0] + "/tests")
from trace_io import read_trace  # noqa: E402


def write_input(trace_path, out_path, max_ghms=None, min_dots=6, kinds=("l2", "hl")):
    tr = read_trace(trace_path)
    gaps = tr.params["gaps"]
    recs = []
    for ev in tr.events:
        if ev[0] not in kinds:
            continue
        d0, d1 = ev[1].dots
        if len(d0) + len(d1) < min_dots:
            continue
        keys = np.concatenate([(d[:, 0].astype(np.uint32) << 16) | d[:, 1].astype(np.uint32) for d in (d0, d1)]).astype("<u4")
        recs.append((len(d0), len(d1), keys))
        if max_ghms and len(recs) >= max_ghms:
            break
    with open(out_path, "wb") as fh:
        fh.write(struct.pack("<i10i", len(gaps), *(list(gaps) + [0] * (10 - len(gaps)))))
        fh.write(struct.pack("<di", tr.params["q_value"], len(recs)))
        for n0, n1, keys in recs:
            fh.write(struct.pack("<IIQi", n0, n1, 1000000, 2))
            fh.write(keys.tobytes())
    return len(recs)


def write_k2k3_input(trace_path, out_path, max_hl=None, filtered=False, pairs_table=False):
    """Input of tests/emu/emu_k2k3: settings, the flattened dataset, the profiles and, per GHM of the trace
    (level 2 and profile x list), its dot sets, area, level and the reference's packed multiplicons.
    filtered: the profiles carry the Link sets of their multiplicons and a profile GHM expects only the multiplicons the
    reference's search kept (homolog-count filter, format flag 0x100); pairs_table: a table_type=pairs trace with its
    pair sets (flag 0x200), so that profile GHMs can be emulated there too."""
    from trace_io import kept_packed
    tr = read_trace(trace_path)
    p, ds = tr.params, tr.dataset
    gaps = p["gaps"]
    sizes = np.diff(ds.list_off).astype(np.int64)
    masked = [np.zeros(int(n), bool) for n in sizes]
    profiles, recs = [], []
    n_hl = 0
    for ev in tr.events:
        if ev[0] == "l2":
            g = ev[1]
            nx, ny = int(sizes[g.x_list]), int(sizes[g.y_list])
            area = (nx * (nx - 1) // 2 if g.x_list == g.y_list else nx * ny) & 0xffffffff   # GHM.cpp:221-236, 32-bit
            recs.append((-1, g.x_list, g.y_list, g.x_list, 2, area, g))
        elif ev[0] == "mask":
            masked[ev[1]][ev[2]:ev[3] + 1] = True
        elif ev[0] == "profile":
            profiles.append(ev[1])
        elif ev[0] == "hl" and (max_hl is None or n_hl < max_hl):
            g, pr = ev[1], profiles[-1]
            area = (pr.L * (int(sizes[g.y_list]) - int(masked[g.y_list].sum()))) & 0xffffffff
            recs.append((len(profiles) - 1, -1, g.y_list, pr.x_object_id, pr.S + 1, area, g))
            n_hl += 1
    with open(out_path, "wb") as fh:
        fmt = (0x100 if filtered else 0) | (0x200 if pairs_table else 0)
        fh.write(struct.pack("<i10i", len(gaps) | fmt, *(list(gaps) + [0] * (10 - len(gaps)))))
        fh.write(struct.pack("<iiiidd", p["gap_size"], p["cluster_gap"], p["anchor_points"], p["mult_hyp"], p["q_value"], p["prob_cutoff"]))
        fh.write(struct.pack("<i", ds.n_lists))
        fh.write(np.ascontiguousarray(ds.list_off, "<i8").tobytes())
        for a in (ds.fam, ds.gid, ds.num_id):
            fh.write(np.ascontiguousarray(a, "<i4").tobytes())
        fh.write(np.ascontiguousarray(ds.flags, "u1").tobytes())
        if pairs_table:
            fh.write(struct.pack("<i", len(ds.pair_ptr) - 1))
            fh.write(np.ascontiguousarray(ds.pair_ptr, "<i8").tobytes())
            fh.write(np.ascontiguousarray(ds.pair_ids, "<i4").tobytes())
        fh.write(struct.pack("<i", len(profiles)))
        for pr in profiles:
            fh.write(struct.pack("<ii", pr.S, pr.L))
            for a in (pr.fam, pr.gid, pr.num_id):
                fh.write(np.ascontiguousarray(a, "<i4").tobytes())
            fh.write(np.ascontiguousarray(pr.flags, "u1").tobytes())
            if filtered:
                fh.write(struct.pack("<i", len(pr.links)))
                fh.write(np.ascontiguousarray(pr.links, "<i4").tobytes())
        fh.write(struct.pack("<i", len(recs)))
        for prof, xl, yl, xobj, level, area, g in recs:
            d0, d1 = g.dots
            keys = np.concatenate([(d[:, 0].astype(np.uint32) << 16) | d[:, 1].astype(np.uint32) for d in (d0, d1)]).astype("<u4")
            fh.write(struct.pack("<iiiiiQII", prof, xl, yl, xobj, level, area, len(d0), len(d1)))
            fh.write(keys.tobytes())
            expected = kept_packed([g]) if (filtered and prof >= 0) else g.packed
            fh.write(struct.pack("<i", len(expected)))
            fh.write(expected)
    return len(recs)


def write_k2k3_synth_input(out_path, n_pairs=6):
    """emu_k2k3 input without a reference trace: the heaviest list pairs of a config-3-shaped set (GHMs with
    100+ clusters: the cooperative merge / insert / statistics paths), dots and expected bytes from the oracle."""
    sys.path.insert(0, __file__.rsplit("/", 3)[0] + "/i-adhore_b200/python")
    import parity_lib as pl
    from iadhore_b200 import capi, ingest, synth
    cfg = synth.config_c3(1.0)
    cfg.genomes = cfg.genomes[:4]
    ds = ingest.dataset_from_synth(cfg)
    eng = pl.oracle_engine(ds)
    eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    n = ds.n_lists
    pairs = [(x, y) for x in range(n) for y in range(x + 1)][::7][:150]
    dots = eng.build_dots([p[0] for p in pairs], [p[1] for p in pairs])
    order = sorted(range(len(pairs)), key=lambda i: -(len(dots[i][0]) + len(dots[i][1])))[:n_pairs]
    gaps = capi.gap_sizes(cfg.gap_size)
    sizes = np.diff(ds.list_off).astype(np.int64)
    with open(out_path, "wb") as fh:
        fh.write(struct.pack("<i10i", len(gaps), *(list(gaps) + [0] * (10 - len(gaps)))))
        fh.write(struct.pack("<iiiidd", cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff))
        fh.write(struct.pack("<i", ds.n_lists))
        fh.write(np.ascontiguousarray(ds.list_off, "<i8").tobytes())
        for a in (ds.fam, ds.gid, ds.num_id):
            fh.write(np.ascontiguousarray(a, "<i4").tobytes())
        fh.write(np.ascontiguousarray(ds.flags, "u1").tobytes())
        fh.write(struct.pack("<i", 0))
        fh.write(struct.pack("<i", len(order)))
        for i in order:
            x, y = pairs[i]
            d0, d1 = dots[i]
            keys = np.concatenate([(d[:, 0].astype(np.uint32) << 16) | d[:, 1].astype(np.uint32) for d in (d0, d1)]).astype("<u4")
            nx, ny = int(sizes[x]), int(sizes[y])
            area = (nx * (nx - 1) // 2 if x == y else nx * ny) & 0xffffffff
            expected = eng.level2_pairs([x], [y]).bytes
            fh.write(struct.pack("<iiiiiQII", -1, x, y, x, 2, area, len(d0), len(d1)))
            fh.write(keys.tobytes())
            fh.write(struct.pack("<i", len(expected)))
            fh.write(expected)
    return len(order)


def write_k1_input(trace_path, out_path):
    """Input of tests/emu/emu_k1: the flat dataset of a (family-mode) trace and, per level-2 GHM, the dot sets the
    reference's buildMatrix produced, as packed keys."""
    tr = read_trace(trace_path)
    ds = tr.dataset
    l2 = [ev[1] for ev in tr.events if ev[0] == "l2"]
    with open(out_path, "wb") as fh:
        fh.write(struct.pack("<i", ds.n_lists))
        fh.write(np.ascontiguousarray(ds.list_off, "<i8").tobytes())
        fh.write(np.ascontiguousarray(ds.fam, "<i4").tobytes())
        fh.write(np.ascontiguousarray(ds.gid, "<i4").tobytes())
        fh.write(np.ascontiguousarray(ds.flags & 1, "u1").tobytes())
        fh.write(struct.pack("<i", len(l2)))
        for g in l2:
            d0, d1 = g.dots
            keys = np.concatenate([(d[:, 0].astype(np.uint32) << 16) | d[:, 1].astype(np.uint32) for d in (d0, d1)]).astype("<u4")
            fh.write(struct.pack("<iiII", g.x_list, g.y_list, len(d0), len(d1)))
            fh.write(keys.tobytes())
        # the profile searches: masks applied before each, the aligned profile, its GHMs with the reference's dots
        import parity_lib as pl
        searches = pl.trace_searches(tr)
        fh.write(struct.pack("<i", len(searches)))
        for masks, pr, ghms in searches:
            fh.write(struct.pack("<i", len(masks)))
            for m in masks:
                fh.write(struct.pack("<iii", *m))
            fh.write(struct.pack("<ii", pr.S, pr.L))
            fh.write(np.ascontiguousarray(pr.fam, "<i4").tobytes())
            fh.write(np.ascontiguousarray(pr.gid, "<i4").tobytes())
            fh.write(np.ascontiguousarray(pr.flags, "u1").tobytes())
            fh.write(struct.pack("<i", len(ghms)))
            for g in ghms:
                d0, d1 = g.dots
                keys = np.concatenate([(d[:, 0].astype(np.uint32) << 16) | d[:, 1].astype(np.uint32) for d in (d0, d1)]).astype("<u4")
                fh.write(struct.pack("<iII", g.y_list, len(d0), len(d1)))
                fh.write(keys.tobytes())
    return len(l2), sum(len(g.dots[0]) + len(g.dots[1]) for g in l2)


def write_k1_workload_input(out_path, workload="d2", n_pairs=40, cfg=None):
    """emu_k1 input from a bench workload (or a given config): the list pairs with the most dots among the pairs of the
    largest lists; expected dots from the oracle's build (pinned on the reference's dot sets by the golden traces)."""
    repo = __file__.rsplit("/", 3)[0]
    sys.path.insert(0, repo + "/i-adhore_b200/python")
    sys.path.insert(0, repo)
    import bench
    import parity_lib as pl
    from iadhore_b200 import capi, ingest
    cfg = cfg or bench.workload_config(workload)
    ds = ingest.dataset_from_synth(cfg)
    eng = pl.oracle_engine(ds)
    sizes = np.diff(ds.list_off).astype(np.int64)
    xs, ys = capi.plan_pairs(sizes)
    cand = np.argsort(-(sizes[xs] * sizes[ys]), kind="stable")[:max(60, n_pairs)]
    dots = eng.build_dots(xs[cand], ys[cand])
    order = sorted(range(len(cand)), key=lambda i: -(len(dots[i][0]) + len(dots[i][1])))[:n_pairs]
    with open(out_path, "wb") as fh:
        fh.write(struct.pack("<i", ds.n_lists))
        fh.write(np.ascontiguousarray(ds.list_off, "<i8").tobytes())
        fh.write(np.ascontiguousarray(ds.fam, "<i4").tobytes())
        fh.write(np.ascontiguousarray(ds.gid, "<i4").tobytes())
        fh.write(np.ascontiguousarray(ds.flags & 1, "u1").tobytes())
        fh.write(struct.pack("<i", len(order)))
        for i in order:
            d0, d1 = dots[i]
            keys = np.concatenate([(d[:, 0].astype(np.uint32) << 16) | d[:, 1].astype(np.uint32) for d in (d0, d1)]).astype("<u4")
            fh.write(struct.pack("<iiII", int(xs[cand[i]]), int(ys[cand[i]]), len(d0), len(d1)))
            fh.write(keys.tobytes())
    return [(len(dots[i][0]), len(dots[i][1])) for i in order]


def write_k1_giant_input(out_path):
    """emu_k1 input with family sizes that take every path of the bucket sort of the join tables -- 1 ... 32 members (one
    register per lane), hundreds (8 / 32 registers), 3,000 (shared memory), 10,000 (global-memory network) -- the
    dataset of tests/test_gpu_parity.py::test_join_tables_with_giant_families; expected dots from the oracle."""
    repo = __file__.rsplit("/", 3)[0]
    sys.path.insert(0, repo + "/i-adhore_b200/python")
    import parity_lib as pl
    from iadhore_b200 import capi
    rng = np.random.RandomState(17)
    sizes = [3000, 10000] + [700, 300, 150, 40] * 3 + list(rng.randint(1, 33, 400))
    fam = np.repeat(np.arange(len(sizes), dtype=np.int32), sizes)
    rng.shuffle(fam)
    n, n_lists = len(fam), 12
    off = np.linspace(0, n, n_lists + 1).astype(np.int64)
    gid = np.arange(n, dtype=np.int32)
    flags = rng.randint(0, 2, n).astype(np.uint8)
    cpu = pl.oracle_engine(capi.Dataset(off, fam, gid, np.arange(n, dtype=np.int32) + 7, flags))
    xs, ys = [0, 3, 11], [0, 4, 1]
    dots = cpu.build_dots(xs, ys)
    with open(out_path, "wb") as fh:
        fh.write(struct.pack("<i", n_lists))
        fh.write(off.astype("<i8").tobytes())
        fh.write(fam.astype("<i4").tobytes())
        fh.write(gid.astype("<i4").tobytes())
        fh.write((flags & 1).astype("u1").tobytes())
        fh.write(struct.pack("<i", len(xs)))
        for x, y, (d0, d1) in zip(xs, ys, dots):
            keys = np.concatenate([(d[:, 0].astype(np.uint32) << 16) | d[:, 1].astype(np.uint32) for d in (d0, d1)]).astype("<u4")
            fh.write(struct.pack("<iiII", x, y, len(d0), len(d1)))
            fh.write(keys.tobytes())
    return sum(len(a) + len(b) for a, b in dots)


def write_tandem_input(trace_path, out_path):
    """Input of tests/emu/emu_tandem: the gene lists of a trace BEFORE tandem remapping (every element: family, gene id,
    strand) with the reference's own Gene::isRemapped() flags as the expected output; pairs traces add the pair sets."""
    tr = read_trace(trace_path)
    raw = tr.raw_lists
    off = np.zeros(len(raw) + 1, np.int64)
    off[1:] = np.cumsum([len(a) for a in raw])
    cat = np.concatenate(raw)
    pairs = hasattr(tr.dataset, "pair_ptr") and tr.dataset.pair_ptr is not None
    with open(out_path, "wb") as fh:
        fh.write(struct.pack("<iii", tr.params["tandem_gap"], 1 if pairs else 0, len(raw)))
        fh.write(off.astype("<i8").tobytes())
        fh.write(np.ascontiguousarray(cat["fam"], "<i4").tobytes())
        fh.write(np.ascontiguousarray(cat["gid"], "<i4").tobytes())
        fh.write(np.ascontiguousarray(cat["strand"], "u1").tobytes())
        fh.write(np.ascontiguousarray(cat["remapped"], "u1").tobytes())
        if pairs:
            fh.write(struct.pack("<i", len(tr.dataset.pair_ptr) - 1))
            fh.write(np.ascontiguousarray(tr.dataset.pair_ptr, "<i8").tobytes())
            fh.write(np.ascontiguousarray(tr.dataset.pair_ids, "<i4").tobytes())
    return int(off[-1]), int(cat["remapped"].sum())


def write_k2k3_workload_input(out_path, workload="d2", n_pairs=4):
    """emu_k2k3 input from a bench workload (bench.workload_config): its n_pairs heaviest list pairs by dot count among
    the pairs of the largest lists; dots and expected bytes from the oracle."""
    repo = __file__.rsplit("/", 3)[0]
    sys.path.insert(0, repo + "/i-adhore_b200/python")
    sys.path.insert(0, repo)
    import bench
    import parity_lib as pl
    from iadhore_b200 import capi, ingest
    cfg = bench.workload_config(workload)
    ds = ingest.dataset_from_synth(cfg)
    eng = pl.oracle_engine(ds)
    eng.set_params(cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff)
    sizes = np.diff(ds.list_off).astype(np.int64)
    xs, ys = capi.plan_pairs(sizes)
    cand = np.argsort(-(sizes[xs] * sizes[ys]), kind="stable")[:60]
    dots = eng.build_dots(xs[cand], ys[cand])
    order = sorted(range(len(cand)), key=lambda i: -(len(dots[i][0]) + len(dots[i][1])))[:n_pairs]
    gaps = capi.gap_sizes(cfg.gap_size)
    with open(out_path, "wb") as fh:
        fh.write(struct.pack("<i10i", len(gaps), *(list(gaps) + [0] * (10 - len(gaps)))))
        fh.write(struct.pack("<iiiidd", cfg.gap_size, cfg.cluster_gap, cfg.anchor_points, capi.MH_BONFERRONI, cfg.q_value, cfg.prob_cutoff))
        fh.write(struct.pack("<i", ds.n_lists))
        fh.write(np.ascontiguousarray(ds.list_off, "<i8").tobytes())
        for a in (ds.fam, ds.gid, ds.num_id):
            fh.write(np.ascontiguousarray(a, "<i4").tobytes())
        fh.write(np.ascontiguousarray(ds.flags, "u1").tobytes())
        fh.write(struct.pack("<i", 0))
        fh.write(struct.pack("<i", len(order)))
        for i in order:
            x, y = int(xs[cand[i]]), int(ys[cand[i]])
            d0, d1 = dots[i]
            keys = np.concatenate([(d[:, 0].astype(np.uint32) << 16) | d[:, 1].astype(np.uint32) for d in (d0, d1)]).astype("<u4")
            nx, ny = int(sizes[x]), int(sizes[y])
            area = (nx * (nx - 1) // 2 if x == y else nx * ny) & 0xffffffff
            expected = eng.level2_pairs([x], [y]).bytes
            fh.write(struct.pack("<iiiiiQII", -1, x, y, x, 2, area, len(d0), len(d1)))
            fh.write(keys.tobytes())
            fh.write(struct.pack("<i", len(expected)))
            fh.write(expected)
    return [(int(xs[cand[i]]), int(ys[cand[i]]), len(dots[i][0]), len(dots[i][1])) for i in order]


def write_synth_input(out_path, workload="c3", n_pairs=12):
    """Dot sets of the heaviest list pairs of a BASELINE-shaped synthetic set (dots from the oracle's build):
    GHMs with 100+ clusters per orientation, which the tiny golden traces do not have."""
    sys.path.insert(0, __file__.rsplit("/", 3)[0] + "/i-adhore_b200/python")
    import parity_lib as pl
    from iadhore_b200 import capi, ingest, synth
    cfg = synth.config_c3(1.0) if workload == "c3" else synth.config_c4(1.0)
    cfg.genomes = cfg.genomes[:4]
    ds = ingest.dataset_from_synth(cfg)
    eng = pl.oracle_engine(ds)
    n = ds.n_lists
    pairs = [(x, y) for x in range(n) for y in range(x + 1)][::7][:150]
    dots = eng.build_dots([p[0] for p in pairs], [p[1] for p in pairs])
    order = sorted(range(len(pairs)), key=lambda i: -(len(dots[i][0]) + len(dots[i][1])))[:n_pairs]
    gaps = capi.gap_sizes(cfg.gap_size)
    with open(out_path, "wb") as fh:
        fh.write(struct.pack("<i10i", len(gaps), *(list(gaps) + [0] * (10 - len(gaps)))))
        fh.write(struct.pack("<di", cfg.q_value, len(order)))
        for i in order:
            d0, d1 = dots[i]
            keys = np.concatenate([(d[:, 0].astype(np.uint32) << 16) | d[:, 1].astype(np.uint32) for d in (d0, d1)]).astype("<u4")
            fh.write(struct.pack("<IIQi", len(d0), len(d1), 1000000, 2))
            fh.write(keys.tobytes())
    return len(order)


if __name__ == "__main__":
    if sys.argv[1] == "--k2k3":
        print(write_k2k3_input(sys.argv[2], sys.argv[3]))
    elif sys.argv[1] == "--synth":
        print(write_synth_input(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "c3", int(sys.argv[4]) if len(sys.argv) > 4 else 12))
    else:
        print(write_input(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None))
