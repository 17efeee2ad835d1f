"""bench.py's reference arm on a tiny workload: runs on CPU (oracle/_ref/ref_harness, or the oracle
port when the reference binary is absent) and must print one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

import parity_lib as pl


import pytest


@pytest.mark.parametrize("no_ref", [False, True], ids=["reference-binary", "oracle-port"])
def test_reference_arm_prints_the_contract_line(no_ref):
    env = dict(os.environ)
    if no_ref:
        env["IADH_BENCH_NO_REF"] = "1"   # the fallback used where oracle/_ref was not built
    p = subprocess.run([sys.executable, os.path.join(pl.REPO, "bench.py"), "--impl", "reference", "--workload", "c3s0.05",
                        "--steps", "1", "--warmup", "1", "--profiles", "8", "--ref-profiles", "2"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-800:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "GHM/s"
    assert line["cpu_baseline"]["kind"] == ("port" if no_ref or not os.path.exists(os.path.join(pl.REPO, "oracle", "_ref", "ref_harness")) else "reference")
    assert line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "same mix" in line["cpu_baseline"]["sample"]
    assert line["config"]["workload"].startswith("c3s0.05")
