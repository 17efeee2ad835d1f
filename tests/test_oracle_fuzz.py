"""Oracle against the UNMODIFIED reference on random small datasets with non-default settings.  Needs
oracle/_ref/ref_harness (built from /root/reference by `make -C oracle ref`; present in the authoring
container and on the GPU box, which receives oracle/_ref with the snapshot); skipped elsewhere.  The
committed golden traces all use the shipped defaults (gap 30, cluster gap 35, q 0.75, cut-off 0.01,
3 anchor points); this covers the rest of the parameter space, bit-exactly (rtol 0), the homolog-count filter
of the profile searches included."""
import os
import random
import shutil
import subprocess
import tempfile

import pytest

import parity_lib as pl
from iadhore_b200 import fuzzcfg, synth
from trace_io import read_trace

HARNESS = os.path.join(pl.REPO, "oracle", "_ref", "ref_harness")


@pytest.mark.skipif(not os.path.exists(HARNESS), reason="oracle/_ref/ref_harness not built (needs the reference tree)")
@pytest.mark.parametrize("seed", [3, 11, 29, 47, 71, 97])
def test_oracle_matches_reference_on_random_settings(seed):
    rng = random.Random(seed)
    checked = 0
    for _ in range(3):
        cfg = fuzzcfg.random_config(rng)
        tmp = tempfile.mkdtemp(prefix="fuzz_")
        try:
            ini = synth.write_reference_inputs(cfg, tmp)
            out = os.path.join(tmp, "trace.bin")
            subprocess.run([HARNESS, ini, out, "--max-profiles", "10"], check=True, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL, timeout=300)
            tr = read_trace(out)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        # (filtered: every profile also through the homolog-count filter, against the reference's own keep / drop flags:
        # the minimum number of homologs per segment is the anchor_points setting, 3 ... 7 here)
        errs = pl.replay_trace(pl.oracle_engine(tr.dataset), tr, rtol_p=0.0, filtered=True)
        assert not errs, (fuzzcfg.describe(cfg), errs[:3])
        checked += sum(1 for e in tr.events if e[0] in ("l2", "hl"))
    assert checked > 30
