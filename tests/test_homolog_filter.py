"""The homolog-count filter of a profile search (GHMProfile.cpp:129-132 = Multiplicon::createHomologs, Multiplicon.cpp:352-533:
pruneLinks + addLinks + per-segment counts), SURVEY 8(f).1, behind iadh_profile_search_filtered /
iadh_profile_search_batch_filtered.

Pinned on the reference's own decisions: the golden traces hold, per profile, the Link set its search inherits and, per
profile x list GHM, which multiplicons the reference's own GHMProfile kept (ref_harness.cpp runs every GHM twice, with
minHomologs = 0 and as shipped).  Three layers:
  CPU   the oracle's restatement (oracle_core.cpp: enough_homologs) reproduces every keep / drop decision;
  CPU   the device source (homolog_kernel.cu: link table kernels + filter kernel, between the emulated clustering kernel
        and the emulated writer) reproduces the kept bytes under the thread-block emulation of tests/emu;
  GPU   the product (C-ABI) returns exactly the kept multiplicons, one search at a time and batched, in every K2 form.
"""
import copy
import json
import os
import subprocess

import numpy as np
import pytest

import parity_lib as pl
from iadhore_b200 import capi
from trace_io import kept_packed, read_trace

TRACES = pl.golden_traces()
IDS = [p.split("/")[-1].split(".")[0] for p in TRACES]
EMU_DIR = os.path.join(pl.HERE, "emu")
EMU_K2K3 = os.path.join(pl.REPO, "build", "emu", "emu_k2k3_%d")
CUDA_INC = "/usr/local/cuda/include/cuda_runtime.h"


def _replay_filtered(eng, tr, rtol_p):
    """Every search of the trace through the filtered entry point; -> (mismatches, multiplicons found, kept)."""
    pl.set_params_from(eng, tr.params)
    eng.clear_masks()
    errs, found, kept = [], 0, 0
    for k, (masks, prof, ghms) in enumerate(pl.trace_searches(tr)):
        for (l, b, e) in masks:
            eng.mask_range(l, b, e)
        r = eng.profile_search_filtered(prof)
        want = kept_packed(ghms)
        msg = capi.compare_packed(r.bytes, want, rtol_p)
        if msg:
            errs.append(f"profile {k}: {msg}")
        n_kept = sum(int(g.kept.sum()) for g in ghms)
        if r.n_multiplicons != n_kept or len(r.ghm_index) != n_kept:
            errs.append(f"profile {k}: stats say {r.n_multiplicons} multiplicons / {len(r.ghm_index)} GHM indices, the reference kept {n_kept}")
        found += sum(len(g.kept) for g in ghms)
        kept += n_kept
    return errs, found, kept


# ---- CPU: the oracle against the reference's decisions ------------------------------------------------------------

@pytest.mark.parametrize("path", TRACES, ids=IDS)
def test_oracle_filter_reproduces_the_reference_decisions(path):
    tr = read_trace(path)
    eng = pl.oracle_engine(tr.dataset)
    errs, found, kept = _replay_filtered(eng, tr, 0.0)
    assert not errs, errs[:5]
    assert 0 < kept < found          # the trace exercises both outcomes


def test_oracle_filter_depends_on_the_inherited_links():
    """Negative control: with most of the inherited Links withheld, multiplicons that survive through pruneLinks alone
    are dropped -- the pinned result is not reachable from addLinks (the family test) alone."""
    tr = read_trace([p for p in TRACES if "tiny_4g" in p][0])
    eng = pl.oracle_engine(tr.dataset)
    pl.set_params_from(eng, tr.params)
    changed = 0
    for masks, prof, ghms in pl.trace_searches(tr):
        for (l, b, e) in masks:
            eng.mask_range(l, b, e)
        p2 = copy.copy(prof)
        p2.links = prof.links[:1]
        changed += eng.profile_search_filtered(p2).bytes != kept_packed(ghms)
    assert changed > 0


def test_filter_needs_a_link_set():
    tr = read_trace(TRACES[0])
    eng = pl.oracle_engine(tr.dataset)
    pl.set_params_from(eng, tr.params)
    prof = copy.copy(pl.trace_searches(tr)[0][1])
    prof.links = np.zeros((0, 4), np.int32)
    with pytest.raises(RuntimeError):
        eng.profile_search_filtered(prof)


# ---- CPU: the device source under the thread-block emulation ---------------------------------------------------------

@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(CUDA_INC):
        pytest.skip("CUDA headers not installed")
    subprocess.run(["make", "-C", EMU_DIR, "all"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    import sys
    sys.path.insert(0, EMU_DIR)
    import make_input
    return make_input


# clustering kernel form x block size; pairs traces carry their pair table (partner buckets rebuilt by the harness)
@pytest.mark.parametrize("name,nt,env", [("tiny_bf", 32, {}), ("tiny_4g", 128, {}), ("tiny_none", 256, {}), ("tiny_fdr", 64, {}),
                                         ("tiny_pairs", 32, {}), ("tiny_pairs_fdr", 128, {}),
                                         ("tiny_4g", 128, {"EMU_RES": "1"}), ("tiny_none", 32, {"EMU_PASS": "2"}),
                                         ("tiny_pairs", 32, {"EMU_PASS": "1"}), ("tiny_pairs_fdr", 64, {"EMU_RES": "1"}),
                                         ("tiny_ap5", 256, {}), ("tiny_ap5", 64, {"EMU_RES": "1"})])   # (5 homologs per segment)
def test_emulated_filter_kernels_reproduce_the_kept_bytes(emu, tmp_path, name, nt, env):
    trace = [p for p in TRACES if p.endswith(name + ".trace.gz")][0]
    inp = str(tmp_path / "in.bin")
    n = emu.write_k2k3_input(trace, inp, filtered=True, pairs_table="pairs" in name)
    p = subprocess.run([EMU_K2K3 % nt, inp], capture_output=True, text=True, env=dict(os.environ, **env), timeout=900)
    assert p.returncode in (0, 1), p.stderr[-500:]
    res = json.loads(p.stdout.strip().splitlines()[-1])
    tr = read_trace(trace)
    hl = [e[1] for e in tr.events if e[0] == "hl"]
    l2 = sum(len(capi.parse_packed(e[1].packed)) for e in tr.events if e[0] == "l2")
    assert res["ghms"] == n
    assert res["before_homolog_filter"] == sum(len(g.kept) for g in hl)           # what the clustering kernels found
    assert res["multiplicons"] == l2 + sum(int(g.kept.sum()) for g in hl)         # what the writer emitted
    assert res["mismatching_ghms"] == 0, p.stderr[-800:]


# ---- GPU: the product through the C-ABI ------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("path", TRACES, ids=IDS)
def test_device_filter_returns_what_the_reference_keeps(path):
    tr = read_trace(path)
    eng = capi.load_product(tr.dataset)
    errs, found, kept = _replay_filtered(eng, tr, pl.P_RTOL)
    eng.close()
    assert not errs, errs[:5]
    assert 0 < kept < found


@pytest.mark.gpu
@pytest.mark.parametrize("knobs", [{}, {"IADH_RESIDENT": "1"}, {"IADH_RESIDENT": "0", "IADH_PASS_MIN": "1", "IADH_PASS_RATIO_HL": "0"},
                                   {"IADH_RESIDENT": "0", "IADH_POOL_FACTOR": "1"}],
                         ids=["default", "resident", "pass-per-launch", "pool-retry"])
@pytest.mark.parametrize("name", ["tiny_4g", "tiny_pairs"])
def test_device_filter_in_a_batch_and_in_every_clustering_form(name, knobs, monkeypatch):
    """iadh_profile_search_batch_filtered over all searches of a trace at once (mask epochs, one Link table per profile)
    against the reference's kept multiplicons, with the clustering stage forced onto each of its forms; then the
    profile-sharded form over 3 ranks."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    tr = read_trace([p for p in TRACES if p.endswith(name + ".trace.gz")][0])
    searches = pl.trace_searches(tr)
    eng = capi.load_product(tr.dataset)
    pl.set_params_from(eng, tr.params)
    profs = [s[1] for s in searches]
    masks = [s[0] for s in searches]
    got = eng.profile_search_batch_filtered(profs, masks)
    for k, (r, s) in enumerate(zip(got, searches)):
        assert capi.compare_packed(r.bytes, kept_packed(s[2]), pl.P_RTOL) is None, k
        assert r.n_multiplicons == sum(int(g.kept.sum()) for g in s[2]), k
    for rank in range(3):
        eng.clear_masks()
        part = eng.profile_search_batch_filtered(profs, masks, rank, 3, True)
        for k, r in enumerate(part):
            if k % 3 == rank:
                assert r.bytes == got[k].bytes, (rank, k)
            else:
                assert r.n_multiplicons == 0
    # the unfiltered entry points are untouched by a filtered call before them
    eng.clear_masks()
    plain = eng.profile_search_batch(profs, masks)
    for k, (r, s) in enumerate(zip(plain, searches)):
        assert r.n_multiplicons == sum(len(g.kept) for g in s[2]), k
    eng.close()


@pytest.mark.gpu
def test_device_filter_rejects_bad_link_sets():
    tr = read_trace(TRACES[0])
    eng = capi.load_product(tr.dataset)
    pl.set_params_from(eng, tr.params)
    prof = copy.copy(pl.trace_searches(tr)[0][1])
    prof.links = np.zeros((0, 4), np.int32)
    with pytest.raises(RuntimeError):
        eng.profile_search_filtered(prof)
    prof.links = np.array([[0, prof.S, 1, 2]], np.int32)      # segment index out of range
    with pytest.raises(RuntimeError):
        eng.profile_search_filtered(prof)
    eng.close()
