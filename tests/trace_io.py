"""Reader for the traces written by oracle/ref_harness.cpp (the unmodified reference
driven GHM by GHM).  Layout (little endian):

  "IADHTRC1" (table_type=family) | "IADHTRC2" (table_type=pairs)
  params : i32 n_gaps, i32 gaps[10], i32 gap_size, cluster_gap, tandem_gap, anchor_points,
           mult_hyp, level2_only, f64 q_value, prob_cutoff
  dataset: i32 n_lists; per list: i32 id, i32 genome_id,
             i32 n_all,  n_all  x {i32 fam, gid, num_id, u8 strand, u8 remapped}
             i32 n,      n      x {i32 fam, gid, num_id, u8 strand}          (remapped elements)
           i32 n, i32 index_to_list[n]
           IADHTRC2 only: i32 n_gene_ids; per gene id: i32 k, k x i32 partner gene id (ascending)
  records: i32 tag
     1 L2 GHM   : i32 x_list, y_list; dots; packed
     2 MASK     : i32 list, begin, end
     3 PROFILE  : i32 S, L, x_object_id; S*L x {i32 fam, gid, num_id, u8 flags(bit0 strand, bit1 gap)}
     4 HL GHM   : i32 y_list; dots; packed (before the homolog filter); i32 n, u8 kept[n]
     5 HL DONE  : i32 n_found
     6 LINKS    : i32 n, n x {i32 segment_x, segment_y, gene_x_id, gene_y_id}  (after a PROFILE record: the Link set
                  the search inherits from the profile's multiplicon, Multiplicon.cpp:343-344; set<Link> order)
     0 end
  dots  : per orientation o=0,1: i32 n, n x {i32 x, i32 y}
  packed: i32 n_bytes, bytes = Multiplicon::packMultiplicons
"""
from __future__ import annotations

import os
import struct
import sys
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "i-adhore_b200", "python"))
from iadhore_b200.capi import Dataset, Profile  # noqa: E402


class _R:
    def __init__(self, b: bytes):
        self.b = b
        self.o = 0

    def i32(self) -> int:
        v = struct.unpack_from("<i", self.b, self.o)[0]
        self.o += 4
        return v

    def f64(self) -> float:
        v = struct.unpack_from("<d", self.b, self.o)[0]
        self.o += 8
        return v

    def u8(self) -> int:
        v = self.b[self.o]
        self.o += 1
        return v

    def rec(self, dtype, n):
        a = np.frombuffer(self.b, dtype=dtype, count=n, offset=self.o)
        self.o += a.nbytes
        return a

    def take(self, n) -> bytes:
        v = self.b[self.o:self.o + n]
        self.o += n
        return v


EL_ALL = np.dtype([("fam", "<i4"), ("gid", "<i4"), ("num_id", "<i4"), ("strand", "u1"), ("remapped", "u1")])
EL = np.dtype([("fam", "<i4"), ("gid", "<i4"), ("num_id", "<i4"), ("strand", "u1")])
DOT = np.dtype([("x", "<i4"), ("y", "<i4")])


@dataclass
class GhmRec:
    kind: str           # "l2" or "hl"
    x_list: int         # l2 only
    y_list: int
    dots: Tuple[np.ndarray, np.ndarray]
    packed: bytes
    kept: np.ndarray = None
    profile_idx: int = -1


@dataclass
class Trace:
    params: dict
    dataset: Dataset
    raw_lists: list      # per list: structured array of ALL elements (pre tandem remap)
    genome_id: list
    index_to_list: np.ndarray
    events: list = field(default_factory=list)   # ("l2", GhmRec) | ("mask", l,b,e) | ("profile", Profile) | ("hl", GhmRec) | ("done", n)


def _dots(r: _R):
    out = []
    for _ in range(2):
        n = r.i32()
        a = r.rec(DOT, n)
        out.append(np.stack([a["x"], a["y"]], axis=1).astype(np.int32) if n else np.zeros((0, 2), np.int32))
    return tuple(out)


def read_trace(path: str) -> Trace:
    import gzip
    with (gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb")) as fh:
        r = _R(fh.read())
    magic = r.take(8)
    assert magic in (b"IADHTRC1", b"IADHTRC2")
    n_gaps = r.i32()
    gaps = [r.i32() for _ in range(10)][:n_gaps]
    prm = dict(gaps=gaps, gap_size=r.i32(), cluster_gap=r.i32(), tandem_gap=r.i32(),
               anchor_points=r.i32(), mult_hyp=r.i32(), level2_only=r.i32(), q_value=r.f64(), prob_cutoff=r.f64())
    nl = r.i32()
    off = [0]
    fam, gid, nid, flg, raw, genome = [], [], [], [], [], []
    for l in range(nl):
        lid = r.i32()
        assert lid == l
        genome.append(r.i32())
        raw.append(r.rec(EL_ALL, r.i32()))
        e = r.rec(EL, r.i32())
        fam.append(e["fam"]); gid.append(e["gid"]); nid.append(e["num_id"]); flg.append(e["strand"].astype(np.uint8))
        off.append(off[-1] + len(e))
    ds = Dataset(np.array(off, np.int64), np.concatenate(fam).astype(np.int32), np.concatenate(gid).astype(np.int32),
                 np.concatenate(nid).astype(np.int32), np.concatenate(flg).astype(np.uint8))
    itl = r.rec("<i4", r.i32()).copy()
    if magic == b"IADHTRC2":   # table_type=pairs: pair sets by gene id (DataSet::getGenePairs)
        ng = r.i32()
        ptr = np.zeros(ng + 1, np.int64)
        ids = []
        for g in range(ng):
            k = r.i32()
            ptr[g + 1] = ptr[g] + k
            ids.append(r.rec("<i4", k))
        ds.pair_ptr = ptr
        ds.pair_ids = np.concatenate(ids).astype(np.int32) if ids else np.zeros(0, np.int32)
    tr = Trace(prm, ds, raw, genome, itl)
    pidx = -1
    while True:
        tag = r.i32()
        if tag == 0:
            break
        if tag == 1:
            x, y = r.i32(), r.i32()
            d = _dots(r)
            pk = r.take(r.i32())
            tr.events.append(("l2", GhmRec("l2", x, y, d, pk)))
        elif tag == 2:
            tr.events.append(("mask", r.i32(), r.i32(), r.i32()))
        elif tag == 3:
            S, L, xo = r.i32(), r.i32(), r.i32()
            e = r.rec(np.dtype([("fam", "<i4"), ("gid", "<i4"), ("num_id", "<i4"), ("flags", "u1")]), S * L)
            pidx += 1
            tr.events.append(("profile", Profile(S, L, xo, e["fam"].copy(), e["gid"].copy(), e["num_id"].copy(), e["flags"].copy())))
        elif tag == 4:
            y = r.i32()
            d = _dots(r)
            pk = r.take(r.i32())
            kept = r.rec("u1", r.i32()).copy()
            tr.events.append(("hl", GhmRec("hl", -1, y, d, pk, kept, pidx)))
        elif tag == 5:
            tr.events.append(("done", r.i32()))
        elif tag == 6:
            lk = r.rec("<i4", 4 * r.i32()).reshape(-1, 4).copy()
            kind, prof = tr.events[-1][0], tr.events[-1][1]
            assert kind == "profile"
            prof.links = lk
        else:
            raise ValueError(f"bad tag {tag} at {r.o}")
    return tr


def concat_packed(chunks: List[bytes]) -> bytes:
    """Concatenate per-GHM packMultiplicons buffers into one stream."""
    n = 0
    body = []
    for c in chunks:
        n += struct.unpack_from("<i", c, 0)[0]
        body.append(c[4:])
    return struct.pack("<i", n) + b"".join(body)


def split_packed(b: bytes) -> List[bytes]:
    """The byte range of every multiplicon of a packMultiplicons buffer (33-byte header, per cluster 102 + 17 * n_ap)."""
    n = struct.unpack_from("<i", b, 0)[0]
    o, out = 4, []
    for _ in range(n):
        begin = o
        ncl = struct.unpack_from("<i", b, o + 29)[0]
        o += 33
        for _ in range(ncl):
            o += 102 + 17 * struct.unpack_from("<i", b, o + 98)[0]
        out.append(b[begin:o])
    assert o == len(b)
    return out


def kept_packed(ghms) -> bytes:
    """What the reference's own search returns for the HL GHMs of one profile: the multiplicons that passed the
    homolog-count filter of GHMProfile.cpp:129-132 (`kept` flags of the trace), as one stream."""
    body = []
    for g in ghms:
        ms = split_packed(g.packed)
        assert len(ms) == len(g.kept)
        body += [m for m, k in zip(ms, g.kept) if k]
    return struct.pack("<i", len(body)) + b"".join(body)
