"""The drop-in shim (integration/dropin_shim.cpp) end to end WITHOUT a GPU: the same shim source linked
against the oracle (tests/dropin_oracle/) replaces the reference's two drivers; the program's seven output
tables must be byte-identical to the unmodified reference binary's.  Checks the shim's own logic --
flattening, mask deltas, profile descriptors, unpack + homolog filter, the pairs table -- on every CPU run;
the product binary (shim + CUDA library) is compared the same way on a GPU by integration/run_e2e.py.
Needs the reference tree (to build the binaries) or binaries built earlier; skipped elsewhere."""
import filecmp
import os
import shutil
import subprocess
import tempfile

import pytest

import parity_lib as pl
from iadhore_b200 import synth

REF = os.path.join(pl.REPO, "oracle", "_ref", "i-adhore")
DROPIN = os.path.join(pl.REPO, "build", "dropin_oracle", "i-adhore-oracle")
TABLES = ["multiplicons.txt", "baseclusters.txt", "anchorpoints.txt", "segments.txt", "list_elements.txt",
          "multiplicon_pairs.txt", "genes.txt"]


def _binaries():
    if os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", os.path.join(pl.REPO, "oracle"), "liboracle.so", "ref"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.run(["make", "-C", os.path.join(pl.REPO, "tests", "dropin_oracle")], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return os.path.exists(REF) and os.path.exists(DROPIN)


def _cfg(name):
    if name == "tiny":
        return synth.config_tiny(seed=11)
    if name == "tiny_fdr":
        return synth.config_tiny(seed=12, mult_hyp="FDR")
    if name == "small":    # 273 profile searches, levels up to 6: exercises the shim's homolog prefilter
        return synth.config_small(seed=21)
    return synth.config_tiny_pairs(seed=15)


# filter: where the homolog-count filter of GHMProfile.cpp:129-132 runs -- "library" = inside the search call
# (iadh_profile_search_filtered, the default; IADH_SHIM_VERIFY_FILTER re-checks every kept multiplicon on the host),
# "host" = the shim's own filter after the unfiltered call (IADH_SHIM_HOST_FILTER)
@pytest.mark.parametrize("name,filt", [("tiny", "library"), ("tiny_fdr", "library"), ("tiny_pairs", "library"), ("small", "library"),
                                       ("tiny_pairs", "host"), ("small", "host")])
def test_dropin_shim_tables_identical_with_oracle_backend(name, filt, monkeypatch):
    if not _binaries():
        pytest.skip("reference tree / prebuilt binaries not available")
    monkeypatch.setenv("IADH_SHIM_VERIFY_FILTER" if filt == "library" else "IADH_SHIM_HOST_FILTER", "1")
    cfg = _cfg(name)
    data = synth.generate(cfg)
    td = tempfile.mkdtemp(prefix="shim_cpu_")
    try:
        rows = {}
        for tag, binary in (("ref", REF), ("dropin", DROPIN)):
            ini = synth.write_reference_inputs(cfg, os.path.join(td, tag), data=data)
            p = subprocess.run([binary, ini], capture_output=True, text=True, timeout=600)
            assert p.returncode == 0, (tag, p.stdout[-400:], p.stderr[-400:])
        for t in TABLES:
            a, b = os.path.join(td, "ref", "out", t), os.path.join(td, "dropin", "out", t)
            assert os.path.exists(a) and os.path.exists(b), t
            assert filecmp.cmp(a, b, shallow=False), f"{name}: {t} differs"
            rows[t] = sum(1 for _ in open(a)) - 1
        assert rows["multiplicons.txt"] > 10 and rows["anchorpoints.txt"] > 100
    finally:
        shutil.rmtree(td, ignore_errors=True)
