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
from trace_io import Trace, concat_packed, kept_packed, read_trace  # noqa: E402

GOLDEN = os.path.join(HERE, "golden")
P_RTOL = 1e-12  # north_star: p-values within 1e-12 relative; everything else bit-exact


def set_params_from(eng: capi.Engine, p: dict):
    eng.set_params(p["gap_size"], p["cluster_gap"], p["anchor_points"], p["mult_hyp"], p["q_value"],
                   p["prob_cutoff"], p["gaps"])


def replay_trace(eng: capi.Engine, tr: Trace, rtol_p: float, check_dots: bool = True, batch_l2: bool = True,
                 filtered: bool = False) -> List[str]:
    """Returns a list of mismatch descriptions (empty == parity).  filtered: every profile is also searched through
    the `_filtered` entry point and compared with what the reference's own search kept (homolog-count filter)."""
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
            if filtered and prof.links is not None and len(prof.links):
                rf = eng.profile_search_filtered(prof)
                msg = capi.compare_packed(rf.bytes, kept_packed(cur), rtol_p)
                if msg:
                    errs.append(f"HL profile {cur[0].profile_idx if cur else '?'}, homolog filter: {msg}")
                n_kept = sum(int(g.kept.sum()) for g in cur)
                if rf.n_multiplicons != n_kept:
                    errs.append(f"HL profile {cur[0].profile_idx if cur else '?'}: {rf.n_multiplicons} multiplicons in the stats, {n_kept} kept by the reference")
            if len(errs) > 20:
                break
    return errs


def trace_searches(tr: Trace):
    """The profile searches of a trace in order: (masks applied before the search, profile, its HL GHM records)."""
    out, masks, prof, cur = [], [], None, []
    for ev in tr.events:
        if ev[0] == "mask":
            masks.append((ev[1], ev[2], ev[3]))
        elif ev[0] == "profile":
            prof, cur = ev[1], []
        elif ev[0] == "hl":
            cur.append(ev[1])
        elif ev[0] == "done":
            out.append((masks, prof, cur))
            masks = []
    return out


def golden_traces(kind: Optional[str] = None) -> List[str]:
    """kind: None = all, "family" / "pairs" = the traces of that table_type (pairs traces are named *pairs*)."""
    out = sorted(os.path.join(GOLDEN, f) for f in os.listdir(GOLDEN) if f.endswith(".trace.gz"))
    if kind is not None:
        out = [p for p in out if ("pairs" in os.path.basename(p)) == (kind == "pairs")]
    return out


def oracle_engine(ds: capi.Dataset) -> capi.Engine:
    return capi.Engine(capi.ORACLE_LIB, "oracle_", ds)


from iadhore_b200.profiles import simple_profile  # noqa: E402,F401


def stream_digest(streams) -> dict:
    """Known-answer digest of packed multiplicon streams (Multiplicon.cpp:210-264 layout): counts, total bytes, sha1 of the
    bytes, sha1 of the bytes with every cluster's 8-byte random_probability zeroed (everything that must be bit-exact
    under the north_star bar) and the sum of log p over all clusters (the one field with a tolerance)."""
    import hashlib
    import math
    import struct
    full, struct_h = hashlib.sha1(), hashlib.sha1()
    n_mult = n_cl = n_bytes = 0
    logs = []
    for b in streams:
        b = bytes(b)
        full.update(b)
        z = bytearray(b)
        (n,) = struct.unpack_from("<i", b, 0)
        o = 4
        for _ in range(n):
            (ncl,) = struct.unpack_from("<i", b, o + 29)
            o += 33
            for _ in range(ncl):
                (pv,) = struct.unpack_from("<d", b, o)
                logs.append(math.log(max(pv, 1e-300)))
                z[o:o + 8] = b"\0" * 8
                o += 102 + 17 * struct.unpack_from("<i", b, o + 98)[0]
                n_cl += 1
        assert o == len(b)
        struct_h.update(bytes(z))
        n_mult += n
        n_bytes += len(b)
    return dict(multiplicons=n_mult, clusters=n_cl, bytes=n_bytes, sha1=full.hexdigest(), struct_sha1=struct_h.hexdigest(),
                p_log_sum=math.fsum(logs))   # (exactly rounded: no summation noise on top of the tolerance below)


def digest_mismatch(got: dict, want: dict, rtol_p: float = P_RTOL) -> Optional[str]:
    """None if `got` meets the parity bar against the known answer: the same bytes, or at least the same bytes outside
    the p-values and a sum of log p that differs by no more than rtol_p per cluster (every p within rtol_p relative,
    the north_star bar, moves its logarithm by at most rtol_p)."""
    for k in ("multiplicons", "clusters", "bytes"):
        if got[k] != want[k]:
            return f"{k}: {got[k]} != {want[k]}"
    if got["sha1"] == want["sha1"]:
        return None
    if got["struct_sha1"] != want["struct_sha1"]:
        return "a field other than a p-value differs"
    if abs(got["p_log_sum"] - want["p_log_sum"]) > rtol_p * want["clusters"] + 1e-9:
        return f"p-values: sum of log p {got['p_log_sum']!r} vs {want['p_log_sum']!r}"
    return None
