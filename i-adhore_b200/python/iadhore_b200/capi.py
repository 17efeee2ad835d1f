"""ctypes binding of the C-ABI in include/iadhore_b200.h.

The same binding drives the product library (`libiadhore_b200.so`, prefix
``iadh_``) and -- from tests / bench cpu_baseline only -- the oracle
(`oracle/liboracle.so`, prefix ``oracle_``), which exports the identical
interface.  The product path never loads the oracle and has no CPU fallback:
`load_product()` raises if the CUDA library is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
PRODUCT_LIB = os.path.join(REPO, "i-adhore_b200", "csrc", "libiadhore_b200.so")
ORACLE_LIB = os.path.join(REPO, "oracle", "liboracle.so")

F_STRAND, F_GAP, F_MASKED = 1, 2, 4
MH_NONE, MH_BONFERRONI, MH_FDR = 0, 1, 2


class DatasetDesc(C.Structure):
    _fields_ = [("n_lists", C.c_int32), ("list_off", C.POINTER(C.c_int64)),
                ("fam", C.POINTER(C.c_int32)), ("gid", C.POINTER(C.c_int32)),
                ("num_id", C.POINTER(C.c_int32)), ("flags", C.POINTER(C.c_uint8))]


class Params(C.Structure):
    _fields_ = [("n_gaps", C.c_int32), ("gaps", C.c_int32 * 10), ("gap_size", C.c_int32),
                ("cluster_gap", C.c_int32), ("anchor_points", C.c_int32), ("mult_hyp", C.c_int32),
                ("q_value", C.c_double), ("prob_cutoff", C.c_double)]


class ProfileDesc(C.Structure):
    _fields_ = [("n_segments", C.c_int32), ("length", C.c_int32), ("x_object_id", C.c_int32),
                ("fam", C.POINTER(C.c_int32)), ("gid", C.POINTER(C.c_int32)),
                ("num_id", C.POINTER(C.c_int32)), ("flags", C.POINTER(C.c_uint8))]


class Stats(C.Structure):
    _fields_ = [("n_ghm", C.c_int64), ("n_dots", C.c_int64), ("n_multiplicons", C.c_int64),
                ("n_baseclusters", C.c_int64), ("n_anchorpoints", C.c_int64),
                ("build_ms", C.c_double), ("cluster_ms", C.c_double), ("total_ms", C.c_double),
                ("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = ["remap_tandems", "remap_tandems_pairs", "intern_strings", "create", "create_pairs", "destroy", "set_params", "level2_all_pairs", "level2_pairs", "mask_range",
           "clear_masks", "profile_search", "profile_search_batch", "profile_search_filtered", "profile_search_batch_filtered", "comm_unique_id", "comm_init", "comm_destroy",
           "allgather_results", "result_bytes", "result_size", "result_ghm_index", "result_stats",
           "free_result", "trim", "warmup", "event_record", "event_elapsed_ms", "build_dots", "build_only", "plan_lists", "plan_pairs", "omxn", "pow", "last_error", "version"]


def gap_sizes(gap_size: int) -> List[int]:
    """Settings::getGapSizes (Settings.cpp:585-598): <=10 distinct round(3^(1+i*step))."""
    import math
    step = (math.log(gap_size) / math.log(3) - 1) / 9
    out: List[int] = []
    for i in range(10):
        v = int(math.floor(math.pow(3, 1 + i * step) + 0.5))  # C round(): half away from zero (positive)
        if i == 0 or v != out[-1]:
            out.append(v)
    return out


@dataclass
class Dataset:
    """Flattened DataSet (after tandem remap), host side (numpy)."""
    list_off: np.ndarray
    fam: np.ndarray
    gid: np.ndarray
    num_id: np.ndarray
    flags: np.ndarray
    # table_type=pairs: CSR over gene ids (gid -> sorted partner gids); None = family mode
    pair_ptr: Optional[np.ndarray] = None
    pair_ids: Optional[np.ndarray] = None

    @property
    def pairs_mode(self) -> bool:
        return self.pair_ptr is not None

    @property
    def n_lists(self) -> int:
        return len(self.list_off) - 1

    def size(self, l: int) -> int:
        return int(self.list_off[l + 1] - self.list_off[l])


@dataclass
class Profile:
    S: int
    L: int
    x_object_id: int
    fam: np.ndarray
    gid: np.ndarray
    num_id: np.ndarray
    flags: np.ndarray
    links: np.ndarray = None   # (n, 4) int32: segment_x, segment_y, gene_x_id, gene_y_id of the parent multiplicon's Link set


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


class Result:
    """One call's packed multiplicon stream.  The stream stays in the library's (pinned) result
    buffer until `.bytes` is first read; the C result is freed when this object dies."""

    def __init__(self, owner, res, n: int, ghm_index: np.ndarray, stats: dict):
        self._owner, self._res, self._ptr, self._n = owner, res, None, n
        self._bytes = None
        self.ghm_index = ghm_index
        self.stats = stats

    def wait(self):
        """Blocks until the stream is in the library's host buffer (the copy overlaps the calls made since)."""
        if self._ptr is None and self._res is not None:
            n = C.c_size_t()
            self._ptr = self._owner._f("result_bytes")(self._res, C.byref(n))
        return self

    @property
    def bytes(self) -> bytes:
        if self._bytes is None:
            self.wait()
            self._bytes = C.string_at(self._ptr, self._n)
            self._release()
        return self._bytes

    @property
    def nbytes(self) -> int:
        return self._n

    def view(self) -> np.ndarray:
        """Zero-copy uint8 view of the packed stream inside the library's pinned result buffer
        (valid while this object is alive and `.bytes` has not been materialised)."""
        if self._bytes is not None:
            return np.frombuffer(self._bytes, dtype=np.uint8)
        self.wait()
        return np.ctypeslib.as_array(self._ptr, shape=(self._n,))

    @property
    def n_multiplicons(self) -> int:
        return int(self.stats["n_multiplicons"])

    def _release(self):
        if self._res is not None:
            self._owner._f("free_result")(self._res)
            self._res = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass


class Engine:
    """One context of a library exporting the iadhore_b200.h interface."""

    def __init__(self, lib_path: str, prefix: str, ds: Dataset, device: int = 0):
        if not os.path.exists(lib_path):
            raise RuntimeError(f"{lib_path} is missing: build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.lib = C.CDLL(lib_path)
        self.prefix = prefix
        f = lambda n: getattr(self.lib, prefix + n)
        self._f = f
        f("last_error").restype = C.c_char_p
        f("version").restype = C.c_char_p
        f("result_bytes").restype = C.POINTER(C.c_uint8)
        f("result_size").restype = C.c_size_t
        f("result_size").argtypes = [C.c_void_p]
        f("result_ghm_index").restype = C.POINTER(C.c_int32)
        f("destroy").restype = None
        f("free_result").restype = None
        f("free_result").argtypes = [C.c_void_p]
        f("trim").restype = None
        self.ds = ds
        self._keep = [np.ascontiguousarray(ds.list_off, dtype=np.int64),
                      np.ascontiguousarray(ds.fam, dtype=np.int32),
                      np.ascontiguousarray(ds.gid, dtype=np.int32),
                      np.ascontiguousarray(ds.num_id, dtype=np.int32),
                      np.ascontiguousarray(ds.flags, dtype=np.uint8)]
        d = DatasetDesc(ds.n_lists, _p(self._keep[0], C.c_int64), _p(self._keep[1], C.c_int32),
                        _p(self._keep[2], C.c_int32), _p(self._keep[3], C.c_int32),
                        _p(self._keep[4], C.c_uint8))
        self.ctx = C.c_void_p()
        if ds.pairs_mode:
            self._keep += [np.ascontiguousarray(ds.pair_ptr, dtype=np.int64), np.ascontiguousarray(ds.pair_ids, dtype=np.int32)]
            self._chk(f("create_pairs")(C.byref(d), _p(self._keep[5], C.c_int64), _p(self._keep[6], C.c_int32),
                                        C.c_int32(len(self._keep[5]) - 1), C.c_int(device), C.byref(self.ctx)))
        else:
            self._chk(f("create")(C.byref(d), C.c_int(device), C.byref(self.ctx)))

    def _chk(self, rc: int):
        if rc != 0:
            raise RuntimeError(f"{self.prefix}* failed ({rc}): {self._f('last_error')().decode()}")

    def close(self):
        if self.ctx:
            self._f("destroy")(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, gap_size=30, cluster_gap=35, anchor_points=3, mult_hyp=MH_BONFERRONI,
                   q_value=0.75, prob_cutoff=0.01, gaps: Optional[Sequence[int]] = None):
        g = list(gaps) if gaps is not None else gap_sizes(gap_size)
        p = Params()
        p.n_gaps = len(g)
        for i in range(10):
            p.gaps[i] = g[i] if i < len(g) else 0
        p.gap_size, p.cluster_gap, p.anchor_points, p.mult_hyp = gap_size, cluster_gap, anchor_points, mult_hyp
        p.q_value, p.prob_cutoff = q_value, prob_cutoff
        self._chk(self._f("set_params")(self.ctx, C.byref(p)))

    def _take(self, res) -> Result:
        n = self._f("result_size")(res)      # (result_bytes would wait for the copy to the host)
        m = C.c_size_t()
        gp = self._f("result_ghm_index")(res, C.byref(m))
        gi = np.ctypeslib.as_array(gp, shape=(m.value,)).copy() if m.value else np.zeros(0, np.int32)
        st = Stats()
        self._chk(self._f("result_stats")(res, C.byref(st)))
        return Result(self, res, int(n), gi, st.as_dict())

    def level2_all_pairs(self, rank: int = 0, n_ranks: int = 1) -> Result:
        res = C.c_void_p()
        self._chk(self._f("level2_all_pairs")(self.ctx, C.c_int(rank), C.c_int(n_ranks), C.byref(res)))
        return self._take(res)

    def level2_pairs(self, xs: Sequence[int], ys: Sequence[int]) -> Result:
        xa = np.ascontiguousarray(xs, dtype=np.int32)
        ya = np.ascontiguousarray(ys, dtype=np.int32)
        res = C.c_void_p()
        self._chk(self._f("level2_pairs")(self.ctx, _p(xa, C.c_int32), _p(ya, C.c_int32),
                                          C.c_int64(len(xa)), C.byref(res)))
        return self._take(res)

    def mask_range(self, lst: int, begin: int, end: int):
        self._chk(self._f("mask_range")(self.ctx, C.c_int32(lst), C.c_int32(begin), C.c_int32(end)))

    def clear_masks(self):
        self._chk(self._f("clear_masks")(self.ctx))

    def _pdesc(self, p: Profile):
        keep = [np.ascontiguousarray(p.fam, dtype=np.int32), np.ascontiguousarray(p.gid, dtype=np.int32),
                np.ascontiguousarray(p.num_id, dtype=np.int32), np.ascontiguousarray(p.flags, dtype=np.uint8)]
        d = ProfileDesc(p.S, p.L, p.x_object_id, _p(keep[0], C.c_int32), _p(keep[1], C.c_int32),
                        _p(keep[2], C.c_int32), _p(keep[3], C.c_uint8))
        return d, keep

    def profile_search(self, p: Profile, rank: int = 0, n_ranks: int = 1) -> Result:
        d, keep = self._pdesc(p)
        res = C.c_void_p()
        self._chk(self._f("profile_search")(self.ctx, C.byref(d), C.c_int(rank), C.c_int(n_ranks), C.byref(res)))
        return self._take(res)

    def profile_search_batch(self, profiles: Sequence[Profile], masks: Sequence[Sequence[Tuple[int, int, int]]],
                             rank: int = 0, n_ranks: int = 1, shard_profiles: bool = False) -> List[Result]:
        """iadh_profile_search_batch: masks[k] = (list, begin, end) ranges applied before profile k."""
        n = len(profiles)
        descs = (ProfileDesc * max(1, n))()
        keep = []
        for k, p in enumerate(profiles):
            d, kp = self._pdesc(p)
            descs[k] = d
            keep.append(kp)
        flat = [m for ms in masks for m in ms]
        marr = np.ascontiguousarray(np.array(flat, dtype=np.int32).reshape(-1, 3))
        mptr = np.zeros(n + 1, np.int32)
        mptr[1:] = np.cumsum([len(ms) for ms in masks])
        outs = (C.c_void_p * max(1, n))()
        self._chk(self._f("profile_search_batch")(self.ctx, descs, C.c_int32(n), marr.ctypes.data_as(C.c_void_p),
                                                   _p(mptr, C.c_int32), C.c_int(rank), C.c_int(n_ranks),
                                                   C.c_int(1 if shard_profiles else 0), outs))
        return [self._take(C.c_void_p(outs[k])) for k in range(n)]

    def profile_search_filtered(self, p: Profile, rank: int = 0, n_ranks: int = 1) -> Result:
        """iadh_profile_search_filtered: the search with the homolog-count filter of GHMProfile.cpp:129-132 applied
        (p.links = the Link set of the profile's multiplicon)."""
        d, keep = self._pdesc(p)
        lk = np.ascontiguousarray(p.links, dtype=np.int32).reshape(-1, 4)
        res = C.c_void_p()
        self._chk(self._f("profile_search_filtered")(self.ctx, C.byref(d), lk.ctypes.data_as(C.c_void_p), C.c_int64(len(lk)),
                                                      C.c_int(rank), C.c_int(n_ranks), C.byref(res)))
        return self._take(res)

    def profile_search_batch_filtered(self, profiles: Sequence[Profile], masks: Sequence[Sequence[Tuple[int, int, int]]],
                                      rank: int = 0, n_ranks: int = 1, shard_profiles: bool = False) -> List[Result]:
        """iadh_profile_search_batch_filtered: the batch with every profile's homolog-count filter applied."""
        n = len(profiles)
        descs = (ProfileDesc * max(1, n))()
        keep = []
        for k, p in enumerate(profiles):
            d, kp = self._pdesc(p)
            descs[k] = d
            keep.append(kp)
        lks = [np.ascontiguousarray(p.links, dtype=np.int32).reshape(-1, 4) for p in profiles]
        lk = np.ascontiguousarray(np.concatenate(lks)) if lks else np.zeros((0, 4), np.int32)
        lptr = np.zeros(n + 1, np.int64)
        lptr[1:] = np.cumsum([len(a) for a in lks])
        flat = [m for ms in masks for m in ms]
        marr = np.ascontiguousarray(np.array(flat, dtype=np.int32).reshape(-1, 3))
        mptr = np.zeros(n + 1, np.int32)
        mptr[1:] = np.cumsum([len(ms) for ms in masks])
        outs = (C.c_void_p * max(1, n))()
        self._chk(self._f("profile_search_batch_filtered")(self.ctx, descs, C.c_int32(n), lk.ctypes.data_as(C.c_void_p),
                                                            _p(lptr, C.c_int64), marr.ctypes.data_as(C.c_void_p),
                                                            _p(mptr, C.c_int32), C.c_int(rank), C.c_int(n_ranks),
                                                            C.c_int(1 if shard_profiles else 0), outs))
        return [self._take(C.c_void_p(outs[k])) for k in range(n)]

    # multi-GPU exchange (iadh_comm_*): rank 0 makes the id, every rank joins with the same id
    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        self._chk(self._f("comm_unique_id")(buf, C.c_size_t(128)))
        return bytes(buf)

    def comm_init(self, uid: bytes, rank: int, n_ranks: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._chk(self._f("comm_init")(self.ctx, buf, C.c_size_t(128), C.c_int(rank), C.c_int(n_ranks)))

    def allgather(self, results: Sequence[Result]) -> List[Result]:
        """iadh_allgather_results: every rank's streams in rank order, per result."""
        n = len(results)
        mine = (C.c_void_p * max(1, n))(*[r._res for r in results])
        outs = (C.c_void_p * max(1, n))()
        self._chk(self._f("allgather_results")(self.ctx, mine, C.c_int32(n), outs))
        return [self._take(C.c_void_p(outs[k])) for k in range(n)]

    def build_dots(self, xs: Sequence[int], ys: Sequence[int], profile: Optional[Profile] = None):
        """-> list of (dots_o0 [n,2], dots_o1 [n,2]) per GHM."""
        xa = np.ascontiguousarray(xs, dtype=np.int32)
        ya = np.ascontiguousarray(ys, dtype=np.int32)
        cnt = C.POINTER(C.c_int64)()
        dots = C.POINTER(C.c_int32)()
        nd = C.c_int64()
        if profile is not None:
            d, keep = self._pdesc(profile)
            pd = C.byref(d)
        else:
            pd = None
        self._chk(self._f("build_dots")(self.ctx, pd, _p(xa, C.c_int32), _p(ya, C.c_int32), C.c_int64(len(ya)),
                                        C.byref(cnt), C.byref(dots), C.byref(nd)))
        n = len(ya)
        cn = np.ctypeslib.as_array(cnt, shape=(2 * n,)).copy() if n else np.zeros(0, np.int64)
        dd = np.ctypeslib.as_array(dots, shape=(nd.value, 2)).copy() if nd.value else np.zeros((0, 2), np.int32)
        out = []
        k = 0
        for i in range(n):
            a = dd[k:k + cn[2 * i]]
            k += cn[2 * i]
            b = dd[k:k + cn[2 * i + 1]]
            k += cn[2 * i + 1]
            out.append((a, b))
        return out

    def build_only(self, xs: Sequence[int], ys: Sequence[int], profile: Optional[Profile] = None) -> dict:
        """Build kernels alone (dots stay on the device): stats dict with n_dots / build_ms."""
        xa = np.ascontiguousarray(xs, dtype=np.int32)
        ya = np.ascontiguousarray(ys, dtype=np.int32)
        if profile is not None:
            d, keep = self._pdesc(profile)
            pd = C.byref(d)
        else:
            pd = None
        st = Stats()
        self._chk(self._f("build_only")(self.ctx, pd, _p(xa, C.c_int32), _p(ya, C.c_int32), C.c_int64(len(ya)), C.byref(st)))
        return st.as_dict()

    def event_record(self, slot: int):
        self._chk(self._f("event_record")(self.ctx, C.c_int(slot)))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_double()
        self._chk(self._f("event_elapsed_ms")(self.ctx, C.c_int(a), C.c_int(b), C.byref(ms)))
        return ms.value

    def omxn(self, x: np.ndarray, n: np.ndarray) -> np.ndarray:
        xa = np.ascontiguousarray(x, dtype=np.float64)
        na = np.ascontiguousarray(n, dtype=np.int32)
        out = np.zeros_like(xa)
        self._chk(self._f("omxn")(self.ctx, _p(xa, C.c_double), _p(na, C.c_int32), C.c_int64(len(xa)), _p(out, C.c_double)))
        return out

    def pow(self, base: np.ndarray, n: np.ndarray) -> np.ndarray:
        xa = np.ascontiguousarray(base, dtype=np.float64)
        na = np.ascontiguousarray(n, dtype=np.int32)
        out = np.zeros_like(xa)
        self._chk(self._f("pow")(self.ctx, _p(xa, C.c_double), _p(na, C.c_int32), C.c_int64(len(xa)), _p(out, C.c_double)))
        return out

    def version(self) -> str:
        return self._f("version")().decode()


def plan_pairs(sizes: Sequence[int], rank: int = 0, n_ranks: int = 1, lib_path: str = PRODUCT_LIB,
               prefix: str = "iadh_") -> Tuple[np.ndarray, np.ndarray]:
    """Host-only: the L2 pair sequence of (rank, n_ranks) (levelTwo.cpp:98-175)."""
    lib = C.CDLL(lib_path)
    sz = np.ascontiguousarray(sizes, dtype=np.int64)
    n = len(sz)
    xs = np.zeros(n * (n + 1) // 2 + 1, np.int32)
    ys = np.zeros_like(xs)
    cnt = C.c_int64()
    rc = getattr(lib, prefix + "plan_pairs")(_p(sz, C.c_int64), C.c_int32(n), C.c_int(rank), C.c_int(n_ranks),
                                             _p(xs, C.c_int32), _p(ys, C.c_int32), C.byref(cnt))
    if rc:
        raise RuntimeError("plan_pairs failed")
    return xs[:cnt.value].copy(), ys[:cnt.value].copy()


def plan_lists(sizes: Sequence[int], rank: int = 0, n_ranks: int = 1, lib_path: str = PRODUCT_LIB,
               prefix: str = "iadh_") -> np.ndarray:
    """Host-only: this rank's y-lists for a profile search (parallel.cpp:96-156)."""
    lib = C.CDLL(lib_path)
    sz = np.ascontiguousarray(sizes, dtype=np.int64)
    out = np.zeros(len(sz) + 1, np.int32)
    cnt = C.c_int64()
    rc = getattr(lib, prefix + "plan_lists")(_p(sz, C.c_int64), C.c_int32(len(sz)), C.c_int(rank), C.c_int(n_ranks),
                                             _p(out, C.c_int32), C.byref(cnt))
    if rc:
        raise RuntimeError("plan_lists failed")
    return out[:cnt.value].copy()


def remap_tandems(list_off, fam, gid, strand, tandem_gap: int, device: int = 0, lib_path: str = PRODUCT_LIB,
                  prefix: str = "iadh_") -> np.ndarray:
    """GeneList::remapTandemsFamily for every list (elements BEFORE remapping): the `remapped` flag
    per element.  Product library: on the device; oracle library: sequential restatement."""
    if not os.path.exists(lib_path):
        raise RuntimeError(f"{lib_path} is missing: build it first")
    lib = C.CDLL(lib_path)
    lo = np.ascontiguousarray(list_off, dtype=np.int64)
    fa = np.ascontiguousarray(fam, dtype=np.int32)
    gi = np.ascontiguousarray(gid, dtype=np.int32)
    fl = np.ascontiguousarray(strand, dtype=np.uint8)
    d = DatasetDesc(len(lo) - 1, _p(lo, C.c_int64), _p(fa, C.c_int32), _p(gi, C.c_int32), _p(gi, C.c_int32), _p(fl, C.c_uint8))
    out = np.zeros(len(fa), np.uint8)
    rc = getattr(lib, prefix + "remap_tandems")(C.byref(d), C.c_int32(tandem_gap), C.c_int(device), _p(out, C.c_uint8))
    if rc:
        f = getattr(lib, prefix + "last_error")
        f.restype = C.c_char_p
        raise RuntimeError(f"{prefix}remap_tandems failed ({rc}): {f().decode()}")
    return out.astype(bool)


def remap_tandems_pairs(list_off, gid, strand, pair_ptr, pair_ids, tandem_gap: int, device: int = 0, lib_path: str = PRODUCT_LIB,
                        prefix: str = "iadh_") -> np.ndarray:
    """GeneList::remapTandemsPairs for every list (elements BEFORE remapping, pair sets as a CSR over the dense gene
    ids): the `remapped` flag per element."""
    if not os.path.exists(lib_path):
        raise RuntimeError(f"{lib_path} is missing: build it first")
    lib = C.CDLL(lib_path)
    lo = np.ascontiguousarray(list_off, dtype=np.int64)
    gi = np.ascontiguousarray(gid, dtype=np.int32)
    fl = np.ascontiguousarray(strand, dtype=np.uint8)
    pp = np.ascontiguousarray(pair_ptr, dtype=np.int64)
    pi = np.ascontiguousarray(pair_ids, dtype=np.int32)
    d = DatasetDesc(len(lo) - 1, _p(lo, C.c_int64), _p(gi, C.c_int32), _p(gi, C.c_int32), _p(gi, C.c_int32), _p(fl, C.c_uint8))
    out = np.zeros(len(gi), np.uint8)
    rc = getattr(lib, prefix + "remap_tandems_pairs")(C.byref(d), _p(pp, C.c_int64), _p(pi, C.c_int32), C.c_int32(len(pp) - 1),
                                                       C.c_int32(tandem_gap), C.c_int(device), _p(out, C.c_uint8))
    if rc:
        f = getattr(lib, prefix + "last_error")
        f.restype = C.c_char_p
        raise RuntimeError(f"{prefix}remap_tandems_pairs failed ({rc}): {f().decode()}")
    return out.astype(bool)


def intern_strings(strings: Sequence[str], device: int = 0, lib_path: str = PRODUCT_LIB, prefix: str = "iadh_") -> Tuple[np.ndarray, int]:
    """iadh_intern_strings: dense ids in order of first occurrence (equal strings, equal ids) and the number of
    distinct strings.  Product library: hash table on the device; oracle library: std::unordered_map."""
    if not os.path.exists(lib_path):
        raise RuntimeError(f"{lib_path} is missing: build it first")
    lib = C.CDLL(lib_path)
    enc = [x.encode() for x in strings]
    off = np.zeros(len(enc) + 1, np.int64)
    if enc:
        off[1:] = np.cumsum([len(b) for b in enc])
    blob = b"".join(enc)
    ids = np.zeros(len(enc), np.int32)
    nd = C.c_int32(0)
    rc = getattr(lib, prefix + "intern_strings")(C.c_char_p(blob), _p(off, C.c_int64), C.c_int64(len(enc)), C.c_int(device),
                                                  _p(ids, C.c_int32), C.byref(nd))
    if rc:
        f = getattr(lib, prefix + "last_error")
        f.restype = C.c_char_p
        raise RuntimeError(f"{prefix}intern_strings failed ({rc}): {f().decode()}")
    return ids, int(nd.value)


def load_product(ds: Dataset, device: int = 0) -> Engine:
    """The B200 path.  Fails loudly when the CUDA library is missing (no fallback)."""
    return Engine(PRODUCT_LIB, "iadh_", ds, device)


# ---------------------------------------------------------------------------
# parsing of the packed multiplicon stream (Multiplicon.cpp:210-264)

@dataclass
class PCluster:
    p: float
    orient: int
    stats: Tuple[float, ...]  # a, b, avg_x, var_x, mrss, x_end1, x_end2, y_end1, y_end2
    was_twisted: int
    bounds: Tuple[int, int, int, int]
    aps: np.ndarray  # [n,5] gx, gy, x, y, real


@dataclass
class PMultiplicon:
    x_object_id: int
    y_object_id: int
    level: int
    redundant: int
    bounds: Tuple[int, int, int, int]
    clusters: List[PCluster]


def parse_packed(b: bytes) -> List[PMultiplicon]:
    off = 0
    (n,) = struct.unpack_from("<i", b, off)
    off += 4
    out = []
    for _ in range(n):
        xo, yo, lv, red, bx, ex, by, ey, nc = struct.unpack_from("<iiiBiiiii", b, off)
        off += 33
        cl = []
        for _ in range(nc):
            p, orient = struct.unpack_from("<dB", b, off)
            off += 9
            st = struct.unpack_from("<9d", b, off)
            off += 72
            tw, cbx, cex, cby, cey, nap = struct.unpack_from("<Biiiii", b, off)
            off += 21
            raw = np.frombuffer(b, dtype=np.uint8, count=17 * nap, offset=off).reshape(nap, 17)
            off += 17 * nap
            ints = raw[:, :16].copy().view("<i4").reshape(nap, 4)
            aps = np.concatenate([ints, raw[:, 16:17].astype(np.int32)], axis=1)
            cl.append(PCluster(p, orient, st, tw, (cbx, cex, cby, cey), aps))
        out.append(PMultiplicon(xo, yo, lv, red, (bx, ex, by, ey), cl))
    assert off == len(b), (off, len(b))
    return out


def compare_packed(got: bytes, want: bytes, rtol_p: float = 1e-12) -> Optional[str]:
    """None if equal: every integer/index/flag field bit-exact, regression doubles
    bit-exact, p-values within rtol_p relative (north_star tolerance)."""
    if got == want:
        return None
    g, w = parse_packed(got), parse_packed(want)
    if len(g) != len(w):
        return f"multiplicon count {len(g)} != {len(w)}"
    for i, (a, b) in enumerate(zip(g, w)):
        if (a.x_object_id, a.y_object_id, a.level, a.redundant, a.bounds) != (b.x_object_id, b.y_object_id, b.level, b.redundant, b.bounds):
            return f"multiplicon {i} header differs: {a} vs {b}"
        if len(a.clusters) != len(b.clusters):
            return f"multiplicon {i}: cluster count {len(a.clusters)} != {len(b.clusters)}"
        for j, (ca, cb) in enumerate(zip(a.clusters, b.clusters)):
            if (ca.orient, ca.was_twisted, ca.bounds) != (cb.orient, cb.was_twisted, cb.bounds):
                return f"multiplicon {i} cluster {j} header differs"
            if ca.aps.shape != cb.aps.shape or not np.array_equal(ca.aps, cb.aps):
                return f"multiplicon {i} cluster {j} anchor points differ"
            sa = struct.pack("<9d", *ca.stats)
            sb = struct.pack("<9d", *cb.stats)
            if sa != sb:
                return f"multiplicon {i} cluster {j} regression stats differ: {ca.stats} vs {cb.stats}"
            if ca.p != cb.p:
                denom = max(abs(cb.p), 1e-300)
                if abs(ca.p - cb.p) / denom > rtol_p:
                    return f"multiplicon {i} cluster {j} p-value {ca.p!r} vs {cb.p!r}"
    return None
