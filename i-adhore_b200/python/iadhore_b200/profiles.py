"""Profile construction helpers for workloads that do not go through the reference's
aligner (bench.py, tests): the hot path (profile x list search) does not depend on which
aligner produced the profile."""
from __future__ import annotations

from typing import List

import numpy as np

from . import capi


def simple_profile(ds: capi.Dataset, xl: int, yl: int, m: capi.PMultiplicon) -> capi.Profile:
    """A 2-segment profile from a level-2 multiplicon, aligned on its anchor points only (the
    hot path does not care which aligner produced the profile; this one is ours: anchor
    points of the first cluster in x order are matched column by column, everything between
    two anchors is laid out x-segment first then y-segment, padded with gaps)."""
    bx, ex, by, ey = m.bounds
    c = m.clusters[0]
    aps = c.aps[np.argsort(c.aps[:, 2], kind="stable")]
    # strictly monotone sub-chain of anchors (in x and, given orientation, in y)
    chain = []
    for gx, gy, x, y, real in aps:
        if not chain:
            chain.append((x, y))
            continue
        px, py = chain[-1]
        if x > px and ((c.orient and y > py) or ((not c.orient) and y < py)):
            chain.append((x, y))
    xo, yo = int(ds.list_off[xl]), int(ds.list_off[yl])
    segx: List[int] = []
    segy: List[int] = []
    yflip = not c.orient

    def emit(xa, xb, ya_list):
        xs_ = list(range(xa, xb))
        n = max(len(xs_), len(ya_list))
        segx.extend(xs_ + [-1] * (n - len(xs_)))
        segy.extend(ya_list + [-1] * (n - len(ya_list)))

    x0, y0 = chain[0]
    segx.append(x0)
    segy.append(y0)
    for (x1, y1) in chain[1:]:
        ys_between = list(range(y0 + 1, y1)) if not yflip else list(range(y0 - 1, y1, -1))
        emit(x0 + 1, x1, ys_between)
        segx.append(x1)
        segy.append(y1)
        x0, y0 = x1, y1
    L = len(segx)
    fam = np.full(2 * L, -1, np.int32)
    gid = np.full(2 * L, -1, np.int32)
    nid = np.full(2 * L, -1, np.int32)
    flg = np.full(2 * L, capi.F_GAP, np.uint8)
    for s, (seg, off, flip) in enumerate(((segx, xo, False), (segy, yo, yflip))):
        for i, e in enumerate(seg):
            if e < 0:
                continue
            k = s * L + i
            fam[k] = ds.fam[off + e]
            gid[k] = ds.gid[off + e]
            nid[k] = ds.num_id[off + e]
            st = int(ds.flags[off + e] & 1)
            flg[k] = (st ^ 1) if flip else st  # inverted y-segments carry flipped strands (GeneList.cpp:270-281)
    return capi.Profile(2, L, xl, fam, gid, nid, flg)


def load_reference_profiles(ds: capi.Dataset, path: str, limit: int = -1, links_path: str = None):
    """The reference's own aligned profiles of a bench workload (tools/make_ref_profiles.py: dumped by
    oracle/_ref/ref_harness --dump-profiles), in the order the reference searches them.
    -> [(masks, Profile)], masks = [(list, begin, end)] applied before that profile (higherLevel.cpp:165-171).
    Elements are identified by ListElement::getNumID(), which the ingest mirror reproduces.
    links_path: the file written by --dump-profile-links beside it; every Profile then carries `.links`, the Link set
    its search inherits (input of the homolog-count filter)."""
    import gzip
    import struct
    raw = gzip.open(path, "rb").read()
    by_num = np.full(int(ds.num_id.max()) + 1, -1, np.int64)
    by_num[ds.num_id] = np.arange(len(ds.num_id))
    out = []
    off = 0
    rec = np.dtype([("num", "<i4"), ("o", "u1")])
    while off < len(raw) and (limit < 0 or len(out) < limit):
        (nm,) = struct.unpack_from("<i", raw, off)
        off += 4
        masks = []
        for _ in range(nm):
            masks.append(struct.unpack_from("<iii", raw, off))
            off += 12
        S, L, xid = struct.unpack_from("<iii", raw, off)
        off += 12
        e = np.frombuffer(raw, dtype=rec, count=S * L, offset=off)
        off += 5 * S * L
        gap = e["num"] < 0
        idx = by_num[np.where(gap, 0, e["num"])]
        if np.any(idx[~gap] < 0):
            raise ValueError("profile element not in the dataset (numID mismatch)")
        fam = np.where(gap, -1, ds.fam[idx]).astype(np.int32)
        gid = np.where(gap, -1, ds.gid[idx]).astype(np.int32)
        nid = np.where(gap, -1, e["num"]).astype(np.int32)
        flg = np.where(gap, capi.F_GAP, e["o"] & 1).astype(np.uint8)
        out.append((masks, capi.Profile(S, L, xid, fam, gid, nid, flg)))
    if links_path:
        import lzma
        lraw = (lzma.open if links_path.endswith(".xz") else gzip.open)(links_path, "rb").read()
        lo = 0
        for _, p in out:
            (nl,) = struct.unpack_from("<i", lraw, lo)
            p.links = np.frombuffer(lraw, dtype="<i4", count=4 * nl, offset=lo + 4).reshape(nl, 4).copy()
            lo += 4 + 16 * nl
    return out
