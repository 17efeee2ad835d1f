"""The C-ABI boundary without a GPU: the product library loads, exports every symbol that
include/iadhore_b200.h declares, its host-only planners agree with the oracle's restatement
of the reference schedulers, and creating a context without a CUDA device fails loudly (no
CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import parity_lib as pl
from iadhore_b200 import capi
from trace_io import read_trace

HEADER = os.path.join(pl.REPO, "include", "iadhore_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(iadh_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = C.CDLL(capi.PRODUCT_LIB)
    syms = declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted("iadh_" + e for e in capi.EXPORTS) == syms


def test_oracle_exports_same_interface():
    lib = C.CDLL(capi.ORACLE_LIB)
    for e in capi.EXPORTS:
        assert hasattr(lib, "oracle_" + e)


def test_struct_sizes_match_header():
    # iadh_params: 15 int32 (+pad) + 2 double; iadh_stats: 5 int64 + 3 double + 3 int64
    assert C.sizeof(capi.Params) == 80
    assert C.sizeof(capi.Stats) == 88
    assert C.sizeof(capi.DatasetDesc) == 48
    assert C.sizeof(capi.ProfileDesc) == 48


@pytest.mark.parametrize("n_ranks", [1, 2, 3, 8])
def test_planners_match_oracle(n_ranks):
    rng = np.random.RandomState(3)
    # (the last one: the 420 ragged list lengths of the datasetII-shaped workload, 127 distinct sizes -> many ties)
    from iadhore_b200 import synth
    d2 = [n for g in synth.config_d2(1.0).genomes for n in g.list_sizes]
    for sizes in ([7], [5, 5, 5], rng.randint(1, 4000, 23).tolist(), [10] * 6 + [3000, 2999, 1], d2):
        allp = set()
        alll = []
        for r in range(n_ranks):
            xa, ya = capi.plan_pairs(sizes, r, n_ranks)
            xb, yb = capi.plan_pairs(sizes, r, n_ranks, capi.ORACLE_LIB, "oracle_")
            assert np.array_equal(xa, xb) and np.array_equal(ya, yb)
            allp |= set(zip(xa.tolist(), ya.tolist()))
            la = capi.plan_lists(sizes, r, n_ranks)
            lb = capi.plan_lists(sizes, r, n_ranks, capi.ORACLE_LIB, "oracle_")
            assert np.array_equal(la, lb)
            alll += la.tolist()
        n = len(sizes)
        assert len(allp) == n * (n + 1) // 2      # every unordered pair exactly once
        assert sorted(alll) == list(range(n))     # every list exactly once


def test_pair_order_matches_reference_trace():
    tr = read_trace(pl.golden_traces()[0])
    sizes = np.diff(tr.dataset.list_off)
    xs, ys = capi.plan_pairs(sizes)
    ref = [(ev[1].x_list, ev[1].y_list) for ev in tr.events if ev[0] == "l2"]
    assert list(zip(xs.tolist(), ys.tolist())) == ref
    assert capi.plan_lists(sizes).tolist() == tr.index_to_list.tolist()


def test_no_cpu_fallback():
    """Without a CUDA device iadh_create must fail with a clear message, never compute."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    tr = read_trace(pl.golden_traces()[0])
    with pytest.raises(RuntimeError, match="no CUDA device|CPU fallback|failed"):
        capi.load_product(tr.dataset)


def test_oracle_interning_is_first_occurrence_order():
    names = ["fam7", "fam2", "fam7", "", "fam2", "x" * 300, "", "fam7a", "fam7"]
    ids, nd = capi.intern_strings(names, lib_path=capi.ORACLE_LIB, prefix="oracle_")
    assert ids.tolist() == [0, 1, 0, 2, 1, 3, 2, 4, 0] and nd == 5
    ids, nd = capi.intern_strings([], lib_path=capi.ORACLE_LIB, prefix="oracle_")
    assert len(ids) == 0 and nd == 0
