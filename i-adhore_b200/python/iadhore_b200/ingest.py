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


def build_dataset(lists: List[List[Gene]], tandem_gap: int, device=None) -> Dataset:
    """lists in ini order (list index == GeneList::getID()).  numID counts every gene of
    every list in order (DataSet.cpp:76-81); gene-ID strings are interned densely.
    device = CUDA ordinal: the tandem remapping runs on the GPU (iadh_remap_tandems, the product
    path); None: the pure-Python mirror above (CPU tests)."""
    gid_map: Dict[str, int] = {}
    off = [0]
    fam_o, gid_o, num_o, flg_o = [], [], [], []
    cnt = 0
    rem_all = None
    if device is not None:
        from . import capi
        raw_off = np.concatenate([[0], np.cumsum([len(l) for l in lists])]).astype(np.int64)
        gids = [gid_map.setdefault(g[0], len(gid_map)) for lst in lists for g in lst]
        rem_all = capi.remap_tandems(raw_off, np.fromiter((g[2] for lst in lists for g in lst), np.int32, int(raw_off[-1])),
                                     np.asarray(gids, np.int32),
                                     np.fromiter((g[1] for lst in lists for g in lst), np.uint8, int(raw_off[-1])),
                                     tandem_gap, device=device)
    for li, lst in enumerate(lists):
        fam = [g[2] for g in lst]
        strand = [g[1] for g in lst]
        gid = [gid_map.setdefault(g[0], len(gid_map)) for g in lst]
        rem = (rem_all[raw_off[li]:raw_off[li + 1]] if rem_all is not None
               else remap_tandems_family(fam, gid, strand, tandem_gap))
        keep = np.nonzero(~rem)[0]
        fam_o.append(np.asarray(fam, np.int32)[keep])
        gid_o.append(np.asarray(gid, np.int32)[keep])
        num_o.append((cnt + keep).astype(np.int32))
        flg_o.append(np.asarray(strand, np.uint8)[keep])
        cnt += len(lst)
        off.append(off[-1] + len(keep))
    cat = lambda xs, dt: (np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt))
    return Dataset(np.asarray(off, np.int64), cat(fam_o, np.int32), cat(gid_o, np.int32), cat(num_o, np.int32),
                   cat(flg_o, np.uint8))


def dataset_from_synth(cfg, data=None, device=None) -> Dataset:
    from . import synth
    if data is None:
        data = synth.generate(cfg)
    lists = [lst for g in cfg.genomes for lst in data[g.name]]
    tg = cfg.tandem_gap if cfg.tandem_gap is not None else cfg.gap_size // 2  # Settings.cpp:337-343
    return build_dataset(lists, tg, device=device)
