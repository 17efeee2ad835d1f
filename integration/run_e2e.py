#!/usr/bin/env python
"""End-to-end drop-in check: the unmodified reference binary (oracle/_ref/i-adhore, CPU) and the
drop-in binary (build/dropin/i-adhore-b200 = the same reference host code + dropin_shim.cpp + the
CUDA library) run the same .ini; every output table must be byte-identical.  Prints wall times.

    python integration/run_e2e.py [--workload small|smallp|c3s0.3|c3] [--threads N] [--keep DIR]

Test/measurement infrastructure: executes oracle/_ref as the checker and CPU baseline.
"""
import argparse
import filecmp
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python"))
from iadhore_b200 import synth  # noqa: E402

REF = os.path.join(REPO, "oracle", "_ref", "i-adhore")
DROPIN = os.path.join(REPO, "build", "dropin", "i-adhore-b200")
TABLES = ["multiplicons.txt", "baseclusters.txt", "anchorpoints.txt", "segments.txt", "list_elements.txt",
          "multiplicon_pairs.txt", "genes.txt"]


def workload(name):
    if name == "small":
        return synth.config_small(seed=21)
    if name == "tiny":
        return synth.config_tiny(seed=11)
    if name in ("tinyp", "smallp"):   # table_type=pairs
        cfg = synth.config_tiny_pairs(seed=15) if name == "tinyp" else synth.config_small(seed=21)
        cfg.table_type = "pairs"
        return cfg
    if name == "c3":
        return synth.config_c3(1.0)
    if name.startswith("c3s"):
        return synth.config_c3(float(name[3:]))
    raise SystemExit("unknown workload " + name)


def run(binary, ini, env=None):
    t0 = time.time()
    p = subprocess.run([binary, ini], capture_output=True, text=True, env=env)
    wall = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError(f"{binary} failed:\n{p.stdout[-800:]}\n{p.stderr[-800:]}")
    l2 = re.search(r"Level 2 multiplicon detection\.\.\.\s*done\. \(time: ([\d.e+-]+)s\)", p.stdout)
    hl = re.search(r"Profile detection\.\.\.\s*done\. \(time: ([\d.e+-]+)s\)", p.stdout)
    return dict(wall_s=wall, l2_s=float(l2.group(1)) if l2 else None, hl_s=float(hl.group(1)) if hl else None,
                stderr="\n".join(l for l in p.stderr.splitlines() if "[iadh]" in l))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="small")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--keep", default=None)
    ap.add_argument("--skip-ref", action="store_true")
    a = ap.parse_args()
    for b in (REF, DROPIN):
        if not os.path.exists(b):
            raise SystemExit(f"{b} missing (built only where /root/reference exists: make -C oracle ref; make -C integration)")
    cfg = workload(a.workload)
    td = a.keep or tempfile.mkdtemp(prefix="iadh_e2e_")
    os.makedirs(td, exist_ok=True)
    data = synth.generate(cfg)
    res = {}
    for tag, binary, thr in (("ref", REF, a.threads), ("b200", DROPIN, 1)):
        d = os.path.join(td, tag)
        ini = synth.write_reference_inputs(cfg, d, data=data, n_threads=thr)
        if tag == "ref" and a.skip_ref:
            continue
        env = dict(os.environ)
        env["IADH_VERBOSE"] = "1"
        res[tag] = run(binary, ini, env)
    out = {"workload": a.workload, "ref_threads": a.threads, "times": res}
    if not a.skip_ref:
        diffs = []
        for t in TABLES:
            fa, fb = os.path.join(td, "ref", "out", t), os.path.join(td, "b200", "out", t)
            if not (os.path.exists(fa) and os.path.exists(fb)):
                diffs.append(t + " (missing)")
            elif not filecmp.cmp(fa, fb, shallow=False):
                diffs.append(t)
        out["tables_identical"] = not diffs
        out["differing_tables"] = diffs
        out["rows"] = {t: sum(1 for _ in open(os.path.join(td, "b200", "out", t))) - 1 for t in TABLES
                       if os.path.exists(os.path.join(td, "b200", "out", t))}
        if res.get("ref") and res.get("b200"):
            out["speedup_wall"] = res["ref"]["wall_s"] / res["b200"]["wall_s"]
    print(json.dumps(out))
    if not a.keep:
        shutil.rmtree(td, ignore_errors=True)
    sys.exit(0 if out.get("tables_identical", True) else 1)


if __name__ == "__main__":
    main()
