#!/usr/bin/env python
"""Run the gprof-instrumented drop-in binary (build/dropin_pg) on a synthetic workload; gmon.out is
left in gpurun_out/pg/ for `gprof build/dropin_pg/i-adhore-b200 gpurun_out/pg/gmon.out`."""
import os, subprocess, sys, time
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(REPO, "i-adhore_b200", "python")); sys.path.insert(0, os.path.join(REPO, "integration"))
from iadhore_b200 import synth
import run_e2e
cfg = run_e2e.workload(sys.argv[1] if len(sys.argv) > 1 else "c3s0.3")
out = os.path.join(REPO, "gpurun_out", "pg")
os.makedirs(out, exist_ok=True)
ini = synth.write_reference_inputs(cfg, os.path.join(out, "data"), n_threads=1)
t0 = time.time()
p = subprocess.run([os.path.join(REPO, "build", "dropin_pg", "i-adhore-b200"), ini], cwd=out, capture_output=True, text=True,
                   env=dict(os.environ, IADH_VERBOSE="1"))
print("rc", p.returncode, "wall", time.time() - t0)
print("\n".join(l for l in p.stdout.splitlines() if "time:" in l))
print(p.stderr[-600:])
import shutil
shutil.rmtree(os.path.join(out, "data"), ignore_errors=True)
