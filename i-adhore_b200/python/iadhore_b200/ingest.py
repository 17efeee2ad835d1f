"""Host-side ingest: gene lists + family table -> the flattened dataset of the C-ABI.

Mirror of the reference's ingest for family mode, which stays on the host (SURVEY.md
section 8 row a0): DataSet::DataSet numbering (DataSet.cpp:55-84), DataSet::getGeneFamilies
(DataSet.cpp:188-202), GeneList::remapTandemsFamily (GeneList.cpp:143-200) with
Gene::remapTo (Gene.cpp:51-62).  tests/test_ingest.py pins it against the reference
(golden traces hold the reference's own remapped lists).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from .capi import Dataset

Gene = Tuple[str, bool, int]  # (gene id, strand, family)


def remap_tandems_family(fam: Sequence[int], gid: Sequence[int], strand: Sequence[bool], gap: int) -> np.ndarray:
    """GeneList::remapTandemsFamily: returns the `remapped` flag per element.

    Three ordered passes within gap+1 positions: '+' genes onto earlier '+'
    genes (back to front), '-' genes onto later '-' genes (front to back), then any still
    unremapped gene onto any earlier gene.  remapTo() also force-clears the target's flag.
    isPairWith in family mode = same family and a different ID string (Gene.cpp:17-30).
    """
    n = len(fam)
    remapped = np.zeros(n, dtype=bool)

    def pair(i, j):
        return fam[i] == fam[j] and gid[i] != gid[j]

    for i in range(n - 1, 0, -1):
        if not strand[i]:
            continue
        j = i - 1
        while j >= 0 and i - j <= gap + 1:
            if strand[j] and pair(i, j):
                remapped[i] = True
                remapped[j] = False
                break
            j -= 1
    for i in range(0, n - 1):
        if strand[i]:
            continue
        j = i + 1
        while j < n and j - i <= gap + 1:
            if (not strand[j]) and pair(i, j):
                remapped[i] = True
                remapped[j] = False
                break
            j += 1
    for i in range(n - 1, 0, -1):
        if remapped[i]:
            continue
        j = i - 1
        while j >= 0 and i - j <= gap + 1:
            if pair(i, j):
                remapped[i] = True
                remapped[j] = False
                break
            j -= 1
    return remapped


def pair_sets(lists: List[List[Gene]], pairs: Sequence[Tuple[str, str]], gid_map: Dict[str, int]):
    """GenePairs::GenePairs (GenePairs.cpp:3-48: symmetric, A == B dropped) followed by
    DataSet::getGenePairs (DataSet.cpp:144-186: partners absent from the gene lists are removed).
    Returns {gid: set(partner gids)}; a gene without an entry has no pair set (Gene::hasPairs false)."""
    present = {g[0] for lst in lists for g in lst}
    for lst in lists:
        for g in lst:
            gid_map.setdefault(g[0], len(gid_map))
    table: Dict[str, set] = {}
    for a, b in pairs:
        if a == b:
            continue
        table.setdefault(a, set()).add(b)
        table.setdefault(b, set()).add(a)
    return {gid_map[a]: {gid_map[b] for b in bs if b in present} for a, bs in table.items() if a in present}


def remap_tandems_pairs(gid: Sequence[int], strand: Sequence[bool], gap: int, ps: Dict[int, set]) -> np.ndarray:
    """GeneList::remapTandemsPairs (GeneList.cpp:83-141) with Gene::isIndirectPairWith (Gene.cpp:33-49):
    a direct pair, or any common partner.  Same three ordered passes as the family variant; a gene
    without a pair set is skipped as the remapped gene (hasPairs) -- and must not occur as the
    candidate target, where the reference dereferences its null set."""
    n = len(gid)
    remapped = np.zeros(n, dtype=bool)
    none = frozenset()

    def indirect(i, j):
        a, b = ps.get(gid[i]), ps.get(gid[j], none)
        return gid[j] in a or not a.isdisjoint(b)

    for i in range(n - 1, 0, -1):
        if gid[i] not in ps or not strand[i]:
            continue
        j = i - 1
        while j >= 0 and i - j <= gap + 1:
            if strand[j] and indirect(i, j):
                remapped[i] = True
                remapped[j] = False
                break
            j -= 1
    for i in range(0, n - 1):
        if gid[i] not in ps or strand[i]:
            continue
        j = i + 1
        while j < n and j - i <= gap + 1:
            if (not strand[j]) and indirect(i, j):
                remapped[i] = True
                remapped[j] = False
                break
            j += 1
    for i in range(n - 1, 0, -1):
        if gid[i] not in ps or remapped[i]:
            continue
        j = i - 1
        while j >= 0 and i - j <= gap + 1:
            if indirect(i, j):
                remapped[i] = True
                remapped[j] = False
                break
            j -= 1
    return remapped


def pairs_csr(ps: Dict[int, set], n_gid: int):
    ptr = np.zeros(n_gid + 1, np.int64)
    for g, v in ps.items():
        ptr[g + 1] = len(v)
    ptr = np.cumsum(ptr).astype(np.int64)
    ids = np.zeros(int(ptr[-1]), np.int32)
    for g, v in ps.items():
        ids[ptr[g]:ptr[g + 1]] = sorted(v)
    return ptr, ids


def build_dataset(lists: List[List[Gene]], tandem_gap: int, device=None, pairs=None) -> Dataset:
    """lists in ini order (list index == GeneList::getID()).  numID counts every gene of
    every list in order (DataSet.cpp:76-81); gene-ID strings are interned densely.
    device = CUDA ordinal: the tandem remapping runs on the GPU (iadh_remap_tandems, the product
    path); None: the pure-Python mirror above (CPU tests)."""
    gid_map: Dict[str, int] = {}
    off = [0]
    fam_o, gid_o, num_o, flg_o = [], [], [], []
    cnt = 0
    rem_all = None
    ps = pair_sets(lists, pairs, gid_map) if pairs is not None else None
    if device is not None and ps is None:
        from . import capi
        raw_off = np.concatenate([[0], np.cumsum([len(l) for l in lists])]).astype(np.int64)
        # gene-ID strings -> dense ids on the device (iadh_intern_strings: first-occurrence order, i.e. exactly the
        # numbering the dict below would produce); the dict is then only a lookup table for the per-list loop
        names = [g[0] for lst in lists for g in lst]
        gid_arr, _ = capi.intern_strings(names, device=device)
        gids = gid_arr.tolist()
        gid_map.update(zip(names, gids))
        rem_all = capi.remap_tandems(raw_off, np.fromiter((g[2] for lst in lists for g in lst), np.int32, int(raw_off[-1])),
                                     np.asarray(gids, np.int32),
                                     np.fromiter((g[1] for lst in lists for g in lst), np.uint8, int(raw_off[-1])),
                                     tandem_gap, device=device)
    if device is not None and ps is not None:
        # table_type=pairs: GeneList::remapTandemsPairs on the device (iadh_remap_tandems_pairs) over the CSR pair sets
        from . import capi
        raw_off = np.concatenate([[0], np.cumsum([len(l) for l in lists])]).astype(np.int64)
        gids = [gid_map.setdefault(g[0], len(gid_map)) for lst in lists for g in lst]
        pp, pi = pairs_csr(ps, len(gid_map))
        rem_all = capi.remap_tandems_pairs(raw_off, np.asarray(gids, np.int32),
                                           np.fromiter((g[1] for lst in lists for g in lst), np.uint8, int(raw_off[-1])),
                                           pp, pi, tandem_gap, device=device)
    for li, lst in enumerate(lists):
        fam = [g[2] for g in lst]
        strand = [g[1] for g in lst]
        gid = [gid_map.setdefault(g[0], len(gid_map)) for g in lst]
        if rem_all is not None:
            rem = rem_all[raw_off[li]:raw_off[li + 1]]
        elif ps is not None:
            rem = remap_tandems_pairs(gid, strand, tandem_gap, ps)
        else:
            rem = remap_tandems_family(fam, gid, strand, tandem_gap)
        keep = np.nonzero(~rem)[0]
        fam_o.append(np.asarray(fam, np.int32)[keep])
        gid_o.append(np.asarray(gid, np.int32)[keep])
        num_o.append((cnt + keep).astype(np.int32))
        flg_o.append(np.asarray(strand, np.uint8)[keep])
        cnt += len(lst)
        off.append(off[-1] + len(keep))
    cat = lambda xs, dt: (np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt))
    ds = Dataset(np.asarray(off, np.int64), cat(fam_o, np.int32), cat(gid_o, np.int32), cat(num_o, np.int32),
                 cat(flg_o, np.uint8))
    if ps is not None:
        ds.pair_ptr, ds.pair_ids = pairs_csr(ps, len(gid_map))
    return ds


def dataset_from_synth(cfg, data=None, device=None) -> Dataset:
    from . import synth
    if data is None:
        data = synth.generate(cfg)
    lists = [lst for g in cfg.genomes for lst in data[g.name]]
    tg = cfg.tandem_gap if cfg.tandem_gap is not None else cfg.gap_size // 2  # Settings.cpp:337-343
    pairs = synth.pair_table(cfg, data) if cfg.table_type == "pairs" else None
    return build_dataset(lists, tg, device=device, pairs=pairs)
