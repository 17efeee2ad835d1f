"""Shared parity drivers: replay a reference trace, or a synthetic workload, through any
library exporting the include/iadhore_b200.h interface and compare against the expected
stream.  Used by the CPU tests (oracle vs golden) and the GPU tests (CUDA path vs golden /
vs oracle)."""
from __future__ import annotations

import os
import sys
from typing import List, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python"))

from iadhore_b200 import capi  # noqa: E402
from trace_io import Trace, concat_packed, read_trace  # noqa: E402

GOLDEN = os.path.join(HERE, "golden")
P_RTOL = 1e-12  # north_star: p-values within 1e-12 relative; everything else bit-exact


def set_params_from(eng: capi.Engine, p: dict):
    eng.set_params(p["gap_size"], p["cluster_gap"], p["anchor_points"], p["mult_hyp"], p["q_value"],
                   p["prob_cutoff"], p["gaps"])


def replay_trace(eng: capi.Engine, tr: Trace, rtol_p: float, check_dots: bool = True, batch_l2: bool = True) -> List[str]:
    """Returns a list of mismatch descriptions (empty == parity)."""
    errs: List[str] = []
    set_params_from(eng, tr.params)
    eng.clear_masks()
    l2 = [ev[1] for ev in tr.events if ev[0] == "l2"]
    if l2:
        xs = [g.x_list for g in l2]
        ys = [g.y_list for g in l2]
        if check_dots:
            dots = eng.build_dots(xs, ys)
            for g, d in zip(l2, dots):
                if not (np.array_equal(d[0], g.dots[0]) and np.array_equal(d[1], g.dots[1])):
                    errs.append(f"L2 dots differ for pair ({g.x_list},{g.y_list})")
        if batch_l2:
            r = eng.level2_pairs(xs, ys)
            msg = capi.compare_packed(r.bytes, concat_packed([g.packed for g in l2]), rtol_p)
            if msg:
                errs.append("L2 batch: " + msg)
            # the all-pairs entry point walks the same sequence on its own
            r2 = eng.level2_all_pairs()
            if r2.bytes != r.bytes:
                errs.append("level2_all_pairs differs from level2_pairs on the reference's pair order")
        else:
            for g in l2:
                r = eng.level2_pairs([g.x_list], [g.y_list])
                msg = capi.compare_packed(r.bytes, g.packed, rtol_p)
                if msg:
                    errs.append(f"L2 ({g.x_list},{g.y_list}): {msg}")
    prof = None
    cur = []
    for ev in tr.events:
        if ev[0] == "mask":
            eng.mask_range(ev[1], ev[2], ev[3])
        elif ev[0] == "profile":
            prof = ev[1]
            cur = []
        elif ev[0] == "hl":
            cur.append(ev[1])
        elif ev[0] == "done":
            if check_dots and cur:
                dots = eng.build_dots([0] * len(cur), [g.y_list for g in cur], profile=prof)
                for g, d in zip(cur, dots):
                    if not (np.array_equal(d[0], g.dots[0]) and np.array_equal(d[1], g.dots[1])):
                        errs.append(f"HL dots differ: profile {g.profile_idx} list {g.y_list}")
            r = eng.profile_search(prof)
            msg = capi.compare_packed(r.bytes, concat_packed([g.packed for g in cur]), rtol_p)
            if msg:
                errs.append(f"HL profile {cur[0].profile_idx if cur else '?'}: {msg}")
            if len(errs) > 20:
                break
    return errs


def golden_traces() -> List[str]:
    return sorted(os.path.join(GOLDEN, f) for f in os.listdir(GOLDEN) if f.endswith(".trace.gz"))


def oracle_engine(ds: capi.Dataset) -> capi.Engine:
    return capi.Engine(capi.ORACLE_LIB, "oracle_", ds)


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
